"""device time of the per-cell kernels at 256^3 (TV gradient, Adam, fused call), CUDA events"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
pkg = ge.package()
t = pkg.RayMarcherExample(0)
t.InitGrid(256)
t.fill_synthetic(20.0, 0)
t.optimizerInit(); t.zeroGrad()
g = np.zeros(4, np.float32)
def tm(name, fn, n=5):
    fn(); t.sync()
    t.timer_start()
    for _ in range(n): fn()
    print("%-28s %.3f ms" % (name, t.timer_stop_ms() / n))
tm("tv_grad (sync api)", lambda: t.kernel2D_TVGrad())
tm("adam all-zero moments", lambda: t.optimizerStep(0))
# make every moment non-zero: gradient = 1 everywhere
import ctypes as C
t.set_grad(np.ones((256 ** 3, 13), np.float32))
tm("adam dense moments", lambda: t.optimizerStep(1))
tm("percell fused call", lambda: t._ck(t.lib.lrf_train_percell_async(t.h, 3)))
tm("adam_zero_async", lambda: t._ck(t.lib.lrf_adam_zero_async(t.h, 4)))
tm("zero_grad_async", lambda: t.zero_grad_async())
# A/B of the TV kernels on the same box, and their agreement
for sweep in (1, 0):
    t.set_option("tv_sweep", sweep)
    tm("tv_grad tv_sweep=%d" % sweep, lambda: t.kernel2D_TVGrad())
    for ov in (2, 1, 0):
        t.set_option("percell_overlap", ov)
        tm("percell fused tv_sweep=%d overlap=%d" % (sweep, ov), lambda: t._ck(t.lib.lrf_train_percell_async(t.h, 5)), n=10)
t.set_option("tv_sweep", 1); t.set_option("percell_overlap", 1)
for bps in (0, 2, 3, 4, 5, 6):
    t.set_option("tv_blocks_per_sm", bps)
    tm("percell overlap, TV blocks/SM cap %d" % bps, lambda: t._ck(t.lib.lrf_train_percell_async(t.h, 5)), n=10)
t.set_option("tv_blocks_per_sm", 0)
t.fill_synthetic(20.0, 0)
res = {}
for sweep in (1, 0):
    t.set_option("tv_sweep", sweep)
    t.zeroGrad()
    t.kernel2D_TVGrad()
    res[sweep] = t.get_grad()[:, 0].astype(np.float64)
d = np.abs(res[1] - res[0])
print("tv sweep vs tile kernel: max abs diff %.3e, max |g| %.3e, nonzero %d" % (d.max(), np.abs(res[0]).max(), int((res[0] != 0).sum())))
