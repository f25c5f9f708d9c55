"""Small driver for ncu captures: renders a few 800x800 views of the bench workload (256^3, D=20) and,
optionally, one training iteration.  Usage: python tools/prof_render.py [n_views] [--batch] [--train]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import bench

pkg = ge.package()
n_views = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 4
for a in sys.argv[1:]:
    if a.startswith("--workload="):
        bench.select_workload(a.split("=", 1)[1])
t = pkg.RayMarcherExample(0)
t.InitGrid(bench.GS)
grid = None
if bench.FILL_MODE is None:
    grid = bench.lcg_grid(bench.GS, bench.DENSITY_SCALE)
    t.set_grid(grid)
else:
    t.fill_synthetic(bench.DENSITY_SCALE, bench.FILL_MODE)
if bench.SPARSE:
    t.RebuildOctree()
proj_inv, views = bench.cameras(pkg, 0, 1)
t.set_proj_inv(proj_inv)
if "--batch" in sys.argv:      # the bench's step: n_views views in ONE launch (twice: ncu -s 1 -c 1 captures the warm one)
    frames = pkg.DeviceBuffer(t, n_views * bench.W * bench.H * 4)
    for _ in range(2):
        t.render_views_dev_async(views[:n_views], frames.ptr, bench.W, bench.H)
else:
    frame = pkg.DeviceBuffer(t, bench.W * bench.H * 4)
    for k in range(n_views):
        t.set_view_inv(views[(k * 13) % len(views)])
        t.render_dev_async(frame.ptr, bench.W, bench.H)
t.sync()
if "--train" in sys.argv:
    r = bench.train_bench(pkg, t, grid, views, proj_inv, 0, 1, None, lambda x: x, t.sync, iters=2)
    print(r)
print("done")
