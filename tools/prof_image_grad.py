"""ncu driver: a few whole-image backward launches (800x800, 256^3 D=20)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
pkg = ge.package()
gs, wh = 256, 800
t = pkg.RayMarcherExample(0)
t.InitGrid(gs); t.fill_synthetic(20.0, 0); t.zeroGrad()
t.set_proj_inv(pkg.inverse4x4(pkg.perspective(45.0, 1.0, 0.1, 100.0)))
t.set_view_inv(pkg.inverse4x4(pkg.orbit_view(2, 7)))
t.targets_resize(1, wh, wh)
t.target_upload(0, np.random.default_rng(0).random((wh * wh, 4)).astype(np.float32))
for _ in range(3):
    t.render_grad_target_async(0)
t.sync()
print("done")
