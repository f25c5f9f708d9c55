import os, sys
import numpy as np
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge, bench
pkg = ge.package()
t = pkg.RayMarcherExample(0)
t.InitGrid(256); t.SetBoundingBox([0,0,0],[1,1,1]); t.fill_synthetic(20.0, 0)
bench.VIEWS = 100
proj_inv, views = bench.cameras(pkg, 0, 1)
t.set_proj_inv(proj_inv)
frame = pkg.DeviceBuffer(t, 800*800*4)
res = {}
for mode in (1, 0):
    t.set_option("staged_render", mode)
    for k in range(0, 100, 4):
        t.set_view_inv(views[k])
        for rep in range(2):
            t.render_dev_async(frame.ptr, 800, 800)
        t.sync(); t.timer_start()
        for rep in range(3):
            t.render_dev_async(frame.ptr, 800, 800)
        res[(mode, k)] = t.timer_stop_ms() / 3
for k in range(0, 100, 4):
    print("view %2d  staged %.2f  gather %.2f" % (k, res[(1,k)], res[(0,k)]))
