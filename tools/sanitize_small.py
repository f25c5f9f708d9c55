"""small run of every round-2 kernel for compute-sanitizer (memcheck): lean render, sparse render with the brick-distance jump,
packed backward (lane-parallel hit enumeration), whole-image backward, TV z-sweep + Adam."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
pkg = ge.package()
gs, w = 64, 96
t = pkg.RayMarcherExample(0)
t.InitGrid(gs); t.SetBoundingBox([0, 0, 0], [1, 1, 1]); t.fill_synthetic(20.0, 0)
t.set_proj_inv(pkg.inverse4x4(pkg.perspective(45.0, 1.0, 0.1, 100.0)))
img = np.zeros(w * w, np.uint32)
for k in (0, 3):
    t.set_view_inv(pkg.inverse4x4(pkg.orbit_view(k, 7)))
    t.kernel2D_RayMarch(img, w, w)
    t.render_lean_debug(w, w)
t.optimizerInit(); t.zeroGrad()
ref = np.random.default_rng(0).random((w * w, 4)).astype(np.float32)
t.kernel2D_RayMarchGrad(w, w, ref)
rng = np.random.default_rng(1)
n = 512
o = (rng.random((n, 3)) * 3 - 1).astype(np.float32)
p1 = (rng.random((n, 3)) * 0.8 + 0.1).astype(np.float32)
d = p1 - o; d /= np.linalg.norm(d, axis=1, keepdims=True)
t.rays_grad(o, d.astype(np.float32), rng.random((n, 4)).astype(np.float32))
t.kernel2D_TVGrad()
t._ck(t.lib.lrf_train_percell_async(t.h, 0)); t.sync()
t.fill_synthetic(20.0, 2); t.RebuildOctree()
for k in (1, 4):
    t.set_view_inv(pkg.inverse4x4(pkg.orbit_view(k, 7)))
    t.kernel2D_RayMarch(img, w, w)
    t.render_debug(w, w, False, 0)
t.zeroGrad(); t.kernel2D_RayMarchGrad(w, w, ref)
t.rays_grad(o, d.astype(np.float32), rng.random((n, 4)).astype(np.float32))
t.close()
print("sanitize run done")
