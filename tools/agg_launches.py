import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i,r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hdr_i]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0
for r in rows[hdr_i+1:]:
    if len(r) <= vi: continue
    val = float(r[vi].replace(',', '')); unit = r[ui]
    if unit == 'ns': val /= 1e3
    elif unit == 'ms': val *= 1e3
    name = re.sub(r'\(.*', '', r[ki]); agg[name][0] += 1; agg[name][1] += val; tot += val
print("total us %.1f" % tot)
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print("%-62s n=%4d %9.1f us %5.1f%%  avg %7.1f" % (k[:62], n, t, 100*t/tot, t/n))
