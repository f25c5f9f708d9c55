"""L2-persisting window on the density plane (option l2_persist_density): forward render and packed-batch backward with and without it.
Usage: python tools/l2_persist.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge, bench
pkg = ge.package()
t = pkg.RayMarcherExample(0)
t.InitGrid(256); t.SetBoundingBox([0, 0, 0], [1, 1, 1]); t.fill_synthetic(20.0, 0)
bench.VIEWS = 100
proj_inv, views = bench.cameras(pkg, 0, 1)
t.set_proj_inv(proj_inv)
W = H = 800
frame = pkg.DeviceBuffer(t, W * H * 4)
rng = np.random.default_rng(1)
n = 4096
vi_dev = pkg.DeviceBuffer(t, len(views) * 64); vi_dev.upload(np.stack(views).astype(np.float32))
picks = np.stack([rng.integers(0, len(views), n), rng.integers(0, W * H, n)], axis=1).astype(np.uint32)
pk, rf = pkg.DeviceBuffer(t, picks.nbytes), pkg.DeviceBuffer(t, n * 16)
pk.upload(picks); rf.upload(rng.random((n, 4)).astype(np.float32))
od, dd = pkg.DeviceBuffer(t, n * 12), pkg.DeviceBuffer(t, n * 12)
t.zeroGrad()
t._ck(t.lib.lrf_gen_rays_dev_async(t.h, n, vi_dev.ptr, pk.ptr, W, H, od.ptr, dd.ptr))
for pct in (0, 100, 60, 0, 100):
    t.set_option("l2_persist_density", pct)
    for k in range(3):
        t.set_view_inv(views[k]); t.render_dev_async(frame.ptr, W, H)
    t.sync(); t.timer_start()
    for k in range(20):
        t.set_view_inv(views[(k * 5) % 100]); t.render_dev_async(frame.ptr, W, H)
    ms_r = t.timer_stop_ms() / 20
    for _ in range(2):
        t._ck(t.lib.lrf_rays_grad_dev_async(t.h, n, od.ptr, dd.ptr, rf.ptr, 0))
    t.sync(); t.timer_start()
    for _ in range(10):
        t._ck(t.lib.lrf_rays_grad_dev_async(t.h, n, od.ptr, dd.ptr, rf.ptr, 0))
    ms_b = t.timer_stop_ms() / 10
    print("l2_persist_density=%3d %%: render %.3f ms per view, packed backward %.3f ms per 4096 rays" % (pct, ms_r, ms_b))
