import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
pkg = ge.package()
t = pkg.RayMarcherExample(0)
t.InitGrid(256); t.fill_synthetic(20.0, 0); t.zeroGrad()
for _ in range(3):
    t.kernel2D_TVGrad()
print("done")
