"""Top source lines of one kernel by warp-stall samples, from `ncu -i x.ncu-rep --page source --csv` output.
usage: python tools/ncu_source_top.py source.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], newline="", errors="replace")))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
for i, r in enumerate(rows):
    if any("Sampling" in c for c in r) and any(c.strip() == "Source" for c in r):
        hdr = i
        break
if hdr is None:
    print("no source table found; first rows:")
    for r in rows[:5]:
        print(r[:12])
    sys.exit(0)
h = rows[hdr]
col_src = [i for i, c in enumerate(h) if c.strip() == "Source"][0]
col_smp = [i for i, c in enumerate(h) if "Sampling" in c and "All" in c]
col_smp = col_smp[0] if col_smp else [i for i, c in enumerate(h) if "Sampling" in c][0]
col_ins = [i for i, c in enumerate(h) if "Instructions Executed" in c]
col_ins = col_ins[0] if col_ins else None
col_no = [i for i, c in enumerate(h) if c.strip() in ("#", "Line")]
data = []
tot = 0.0
for r in rows[hdr + 1:]:
    if len(r) <= max(col_src, col_smp):
        continue
    try:
        s = float(r[col_smp].replace(",", "") or 0)
    except ValueError:
        continue
    tot += s
    data.append((s, r[col_no[0]] if col_no else "", r[col_src].strip()[:110], r[col_ins] if col_ins is not None else ""))
data.sort(key=lambda t: -t[0])
print("total samples %.0f" % tot)
for s, no, src, ins in data[:n]:
    print("%6.0f %5.1f%%  L%-5s inst %-9s %s" % (s, 100.0 * s / max(tot, 1.0), no, ins, src))
