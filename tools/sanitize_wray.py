"""packed backward alone (for compute-sanitizer --tool racecheck: the rank table of WarpRay::gather_parallel lives in shared memory)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
pkg = ge.package()
t = pkg.RayMarcherExample(0)
t.InitGrid(64); t.fill_synthetic(20.0, 0); t.zeroGrad()
rng = np.random.default_rng(1)
n = 256
o = (rng.random((n, 3)) * 3 - 1).astype(np.float32)
p1 = (rng.random((n, 3)) * 0.8 + 0.1).astype(np.float32)
d = p1 - o; d /= np.linalg.norm(d, axis=1, keepdims=True)
t.rays_grad(o, d.astype(np.float32), rng.random((n, 4)).astype(np.float32))
t.close()
print("done")
