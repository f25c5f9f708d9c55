"""Per-source-line dynamic instruction counts: joins the per-address counts of an .ncu-rep SASS page with the
line table of the cubin inside liblrf_b200.so (nvdisasm --print-line-info).
Usage: python tools/ncu_lines.py gpurun_out/x.ncu-rep <kernel-name-substring> [top_n]"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kname = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "literf-radiancefields_b200", "_build", "liblrf_b200.so")], cwd=tmp, capture_output=True)
cubin = [f for f in glob.glob(os.path.join(tmp, "*.cubin")) if os.path.basename(f).startswith("lrf_kernels.")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
# locate the kernel's text section
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and kname in l][0]
off2line, cur = {}, None
for l in dis[start + 1:]:
    if l.startswith("\t.section") or (l.startswith(".text.") and kname not in l):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*)", l)
    if m and cur:
        off2line[int(m.group(1), 16)] = cur
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hdr]
ie, ist = h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
body = [r for r in rows[hdr + 1:] if r and len(r) > max(ie, ist) and r[ie].isdigit()]
base = int(body[0][0], 16) if body[0][0].startswith("0x") else int(body[0][0])
cnt, st = collections.Counter(), collections.Counter()
tot = 0
for r in body:
    a = int(r[0], 16) if r[0].startswith("0x") else int(r[0])
    key = off2line.get(a - base, ("?", 0))
    cnt[key] += int(r[ie])
    st[key] += int(r[ist]) if r[ist].isdigit() else 0
    tot += int(r[ie])
src = {}
for f in ("lrf_kernels.cu", "lrf_device.cuh", "lrf_fast.cuh", "lrf_wray.cuh"):
    src[f] = open(os.path.join(ROOT, "literf-radiancefields_b200", "csrc", f)).read().splitlines()
ts = max(1, sum(st.values()))
print("total dynamic warp-instr", tot)
order = st.most_common(topn) if "--by-stall" in sys.argv else cnt.most_common(topn)
for (f, ln), _v in order:
    v = cnt[(f, ln)]
    text = src.get(f, [""] * (ln + 1))[ln - 1].strip()[:110] if f in src and ln - 1 < len(src[f]) else ""
    print("%5.2f%% instr %5.2f%% stall  %s:%d  %s" % (100 * v / tot, 100 * st[(f, ln)] / ts, f, ln, text))
