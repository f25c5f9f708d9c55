"""Print the SASS of one kernel of liblrf_b200.so (name substring) — instruction lines only.
Usage: python tools/sass_of.py <substring> [out_file]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "literf-radiancefields_b200", "_build", "liblrf_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
out, on = [], False
for line in txt.splitlines():
    if "Function :" in line:
        on = sys.argv[1] in line
        if on:
            out.append(line.strip())
        continue
    if on and "/*" in line and not line.strip().startswith("/* 0x"):
        out.append(line.split("/* 0x")[0].rstrip())
text = "\n".join(out)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
else:
    print(text)
