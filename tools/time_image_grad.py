"""device time of the whole-image backward (kernel2D_RayMarchGrad) next to the forward render of the same view"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
pkg = ge.package()
for gs, wh in ((128, 256), (256, 800)):
    t = pkg.RayMarcherExample(0)
    t.InitGrid(gs)
    t.fill_synthetic(20.0, 0)
    t.zeroGrad()
    t.set_proj_inv(pkg.inverse4x4(pkg.perspective(45.0, 1.0, 0.1, 100.0)))
    t.set_view_inv(pkg.inverse4x4(pkg.orbit_view(2, 7)))
    t.targets_resize(1, wh, wh)
    t.target_upload(0, np.random.default_rng(0).random((wh * wh, 4)).astype(np.float32))
    frame = pkg.DeviceBuffer(t, wh * wh * 4)
    for name, fn in (("forward", lambda: t.render_dev_async(frame.ptr, wh, wh)), ("backward", lambda: t.render_grad_target_async(0))):
        fn(); t.sync()
        t.timer_start()
        for _ in range(5): fn()
        ms = t.timer_stop_ms() / 5
        print("gs %d  %dx%d  %-8s %.3f ms  (%.1f Mrays/s)" % (gs, wh, wh, name, ms, wh * wh / ms / 1e3))
    t.close()
