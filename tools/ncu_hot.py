"""SASS of a captured kernel with per-instruction execution counts, restricted to instructions executed at least frac * max times.
Usage: python tools/ncu_hot.py x.ncu-rep [frac=0.5] [kernel-index=0]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hh = rows[hdr[which]]
end = hdr[which + 1] - 1 if len(hdr) > which + 1 else len(rows)
body = rows[hdr[which] + 1:end]
ie, isrc, ist = hh.index("Instructions Executed"), hh.index("Source"), hh.index("Warp Stall Sampling (All Samples)")
ith = hh.index("Thread Instructions Executed") if "Thread Instructions Executed" in hh else None
mx = max(int(r[ie]) for r in body if r[ie].isdigit())
n = 0
tot = 0
for r in body:
    if r[ie].isdigit() and int(r[ie]) >= frac * mx:
        n += 1
        tot += int(r[ie])
        lanes = (int(r[ith]) / max(1, int(r[ie]))) if ith is not None and r[ith].isdigit() else 0
        print("%s %9d %5.1f %6s  %s" % (r[0][-5:], int(r[ie]), lanes, r[ist], r[isrc]))
print("instructions listed:", n, "max count", mx, "sum/max %.1f" % (tot / mx))
