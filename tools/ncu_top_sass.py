"""Top stall-sampled SASS instructions of one kernel instance of an .ncu-rep (captured with --import-source on).
usage: ncu_top_sass.py report.ncu-rep launch_index [top_n]"""
import csv, subprocess, sys
rep, idx = sys.argv[1], int(sys.argv[2])
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(idx),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
h = rows[hi]
si, src, ie = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
data = [r for r in rows[hi + 1:] if len(r) > si and r[si].isdigit()]
tot = sum(int(r[si]) for r in data)
print(rows[0][1][:80], "total samples", tot)
top = sorted(enumerate(data), key=lambda x: -int(x[1][si]))[:topn]
for i, r in sorted(top):
    print("%5d %6s %5.1f%% %9s  %s" % (i, r[si], 100.0 * int(r[si]) / max(tot, 1), r[ie], r[src][:120]))
