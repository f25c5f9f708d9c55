"""Register-window march (lrf_lean.cuh) against the FastMarch gather kernel on the bench workload: pixel differences, shaded counts and
final transmittance bits against the exact parity tap (debug_kernel), and same-box timings of both kernels.
Usage: python tools/lean_check.py [n_views] [--gs=256] [--wh=800] [--time=N] [--D=20]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import bench

opt = dict(a.lstrip("-").split("=", 1) for a in sys.argv[1:] if "=" in a)
n_views = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 3
gs, wh, ntime = int(opt.get("gs", 256)), int(opt.get("wh", 800)), int(opt.get("time", 20))
D = float(opt.get("D", 20.0))
pkg = ge.package()
t = pkg.RayMarcherExample(0)
t.InitGrid(gs)
t.SetBoundingBox([0, 0, 0], [1, 1, 1])
t.fill_synthetic(D, 0)
bench.VIEWS = 100
proj_inv, views = bench.cameras(pkg, 0, 1)
t.set_proj_inv(proj_inv)
W = H = wh
bad = 0
for k in range(n_views):
    vi = views[(k * 13 + int(opt.get("v0", 0))) % len(views)]
    t.set_view_inv(vi)
    t.set_option("lean_render", 0)
    ref = np.zeros(W * H, np.uint32)
    t.kernel2D_RayMarch(ref, W, H)
    t.set_option("lean_render", 1)
    img2 = np.zeros(W * H, np.uint32)
    t.kernel2D_RayMarch(img2, W, H)
    img, ns, thr = t.render_lean_debug(W, H)
    ch_r = np.stack([(ref >> s) & 255 for s in (0, 8, 16, 24)]).astype(np.int32)
    ch_g = np.stack([(img >> s) & 255 for s in (0, 8, 16, 24)]).astype(np.int32)
    diff = np.abs(ch_r - ch_g)
    nh_e, ns_e, rg_e, _ = t.render_debug(W, H, False, 0)
    dec = (ns != 0xFFFFFFFF) & (nh_e > 0)
    n_bad = int((ns[dec] != ns_e[dec]).sum())
    thr_bad = int((thr[dec].view(np.uint32) != rg_e[dec, 3].view(np.uint32)).sum())
    same_frame = bool((img == img2).all())
    print("view %3d  pixels: maxdiff %d frac!=0 %.2e | shaded-count mismatches %d, transmittance-bit mismatches %d of %d rays (%d to fallback) | tap frame == plain frame %s"
          % (k, diff.max(), (diff != 0).mean(), n_bad, thr_bad, int(dec.sum()), int((~dec).sum()), same_frame))
    bad += n_bad + thr_bad + (diff.max() > 1) + (not same_frame)
frame = pkg.DeviceBuffer(t, W * H * 4)
for mode in (1, 0, 1, 0):
    t.set_option("lean_render", mode)
    for k in range(3):
        t.set_view_inv(views[k])
        t.render_dev_async(frame.ptr, W, H)
    t.sync()
    t.timer_start()
    for k in range(ntime):
        t.set_view_inv(views[(k * 5) % len(views)])
        t.render_dev_async(frame.ptr, W, H)
    print("lean_render=%d: %.3f ms per view (%d views)" % (mode, t.timer_stop_ms() / ntime, ntime))
print("BAD" if bad else "OK")
sys.exit(1 if bad else 0)
