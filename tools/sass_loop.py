"""Instruction mix of the innermost hot loop of a kernel: python tools/sass_loop.py <kernel substring> [min_ldg]
Finds backward branches in the SASS, picks the span with the most LDG instructions and prints its opcode histogram."""
import os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "literf-radiancefields_b200", "_build", "liblrf_b200.so") if len(sys.argv) < 4 else sys.argv[3]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ins, on = [], False
for line in txt.splitlines():
    if "Function :" in line:
        if on: break
        on = sys.argv[1] in line
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if on and m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA\S*\s+(?:\S+,\s*)?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr:
            loops.append((addr[tgt], i))
best = None
for lo, hi in loops:
    n = sum(1 for _, t in ins[lo:hi + 1] if "LDG" in t)
    print("loop %#x..%#x: %d instr, %d LDG" % (ins[lo][0], ins[hi][0], hi - lo + 1, n))
    if best is None or n > best[0] or (n == best[0] and hi - lo < best[2] - best[1]):
        best = (n, lo, hi)
if best:
    n, lo, hi = best
    h = collections.Counter()
    for _, t in ins[lo:hi + 1]:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        h[t.split()[0].split(".")[0]] += 1
    print("picked %#x..%#x (%d instr)" % (ins[lo][0], ins[hi][0], hi - lo + 1))
    for k, v in h.most_common():
        print("  %-10s %d" % (k, v))
    if "--dump" in sys.argv:
        for a, t in ins[lo:hi + 1]:
            print("%06x  %s" % (a, t))
