"""Aggregates an ncu launch list (--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv)
per kernel: launches, total / average duration, share of the step, DRAM bytes per launch and the bandwidth they imply.
usage: python tools/agg_launches.py launches.csv [top N]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hdr_i]
ki, vi, ui, mi, ii = (hdr.index(c) for c in ("Kernel Name", "Metric Value", "Metric Unit", "Metric Name", "ID"))
SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
agg = collections.defaultdict(lambda: [set(), 0.0, 0.0])          # launch ids, us, dram bytes
tot = 0.0
for r in rows[hdr_i + 1:]:
    if len(r) <= vi:
        continue
    val = float(r[vi].replace(",", "")) * SCALE.get(r[ui], 1.0)
    name = re.sub(r"\(.*", "", r[ki])
    a = agg[name]
    a[0].add(r[ii])
    if r[mi].startswith("gpu__time_duration"):
        a[1] += val
        tot += val
    elif r[mi].startswith("dram__bytes"):
        a[2] += val
n_show = int(sys.argv[2]) if len(sys.argv) > 2 else 45
print("total us %.1f, launches %d" % (tot, sum(len(a[0]) for a in agg.values())))
for k, (ids, t, by) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:n_show]:
    n = len(ids)
    line = "%-58s n=%3d %8.1f us %5.1f%% avg %7.1f us" % (k[:58], n, t, 100 * t / tot, t / n)
    if by > 0:
        line += "  dram %7.1f MB/launch %7.0f GB/s" % (by / n / 1e6, by / t / 1e3)
    print(line)
