"""Staged render (lrf_staged.cuh) against the exact gather kernel on the bench workload: pixel differences, shaded counts of the
staged kernel against the exact parity tap, fallback counters, and same-box timings of both kernels.
Usage: python tools/staged_check.py [n_views] [--gs=256] [--wh=800] [--nw=4|8] [--time=N] [--notap]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import bench

opt = dict(a.lstrip("-").split("=", 1) for a in sys.argv[1:] if "=" in a)
n_views = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 3
gs, wh, nw, ntime = int(opt.get("gs", 256)), int(opt.get("wh", 800)), int(opt.get("nw", 4)), int(opt.get("time", 20))
D = float(opt.get("D", 20.0))
pkg = ge.package()
t = pkg.RayMarcherExample(0)
t.InitGrid(gs)
t.SetBoundingBox([0, 0, 0], [1, 1, 1])
t.fill_synthetic(D, 0)
bench.VIEWS = 100
proj_inv, views = bench.cameras(pkg, 0, 1)
t.set_proj_inv(proj_inv)
t.set_option("staged_warps", nw)
if "bw" in opt:
    t.set_option("staged_box_w", int(opt["bw"]))
if "bh" in opt:
    t.set_option("staged_box_h", int(opt["bh"]))
W = H = wh
bad = 0
for k in range(n_views):
    vi = views[(k * 13 + int(opt.get("v0", 0))) % len(views)]
    t.set_view_inv(vi)
    t.set_option("staged_render", 0)
    ref = np.zeros(W * H, np.uint32)
    t.kernel2D_RayMarch(ref, W, H)
    t.set_option("staged_render", 1)
    t.render_stats(reset=True)
    t0 = time.perf_counter()
    img, ns = t.render_staged_debug(W, H)
    dt = time.perf_counter() - t0
    st = t.render_stats()
    ch_r = np.stack([(ref >> s) & 255 for s in (0, 8, 16, 24)]).astype(np.int32)
    ch_g = np.stack([(img >> s) & 255 for s in (0, 8, 16, 24)]).astype(np.int32)
    diff = np.abs(ch_r - ch_g)
    line = "view %3d staged=%d  maxdiff %d  frac!=0 %.2e  exact-rays: ambiguous %d other %d  tiles-unfit %d  (%.1f ms wall)" % (
        k, st[3], diff.max(), (diff != 0).mean(), st[0], st[1], st[2], dt * 1e3)
    if "--notap" not in sys.argv:
        _, ns_e, _, _ = t.render_debug(W, H, False, 0)
        dec = ns != 0xFFFFFFFF
        nbad = int((ns[dec] != ns_e[dec]).sum())
        line += "  shaded-count mismatches %d of %d decided" % (nbad, int(dec.sum()))
        bad += nbad
    if diff.max() > 1:
        bad += 1
        idx = np.argwhere(diff.max(axis=0) > 1).ravel()[:8]
        line += "  BAD pixels e.g. %s" % [(int(i % W), int(i // W), hex(int(ref[i])), hex(int(img[i]))) for i in idx]
    print(line, flush=True)

frame = pkg.DeviceBuffer(t, W * H * 4)
for mode in (1, 0):
    t.set_option("staged_render", mode)
    for rep in range(2):
        t.sync()
        t.timer_start()
        for k in range(ntime):
            t.set_view_inv(views[(k * 5) % len(views)])
            t.render_dev_async(frame.ptr, W, H)
        ms = t.timer_stop_ms()
    print("%s: %.3f ms/view  %.1f Mrays/s" % ("staged" if mode else "gather", ms / ntime, W * H * ntime / ms / 1e3), flush=True)
print("RESULT", "FAIL" if bad else "OK")
