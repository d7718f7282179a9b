"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (optionally only the last 1/N)."""
import csv, collections, sys
path = sys.argv[1]
frac = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rows = list(csv.reader(open(path)))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr = i; break
h = rows[hdr]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); gi = h.index('Grid Size')
L = []
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    try: v = float(r[vi].replace(',', ''))
    except ValueError: continue
    L.append((r[ki].split('(')[0].split('::')[-1][:40], r[gi], v / 1e3))
L = L[len(L) - len(L) // frac:]
agg = collections.defaultdict(lambda: [0, 0.0])
for n, g, v in L:
    agg[n][0] += 1; agg[n][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-42s %5d %10.1f us %5.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
print("total %.1f us, %d launches" % (tot, sum(v[0] for v in agg.values())))
if len(sys.argv) > 3:
    for n, g, v in L:
        if sys.argv[3] in n: print("%-30s %-14s %8.1f" % (n, g, v))
