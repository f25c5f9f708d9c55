"""One-GPU training iteration at 256^3 (4096-ray batches), two schedules on the same box:
  A  backward, then the per-cell pass {TV -> Adam(density)} beside Adam(colour)          (lrf_rays_grad_dev_async + lrf_train_percell_async)
  B  backward beside the TV gradient (atomic adds), then Adam + zeroing over all planes   (lrf_rays_grad_tv_dev_async + lrf_adam_zero_async)
CUDA events on the context stream."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import bench
pkg = ge.package()
t = pkg.RayMarcherExample(0)
t.InitGrid(bench.GS)
t.SetBoundingBox([0, 0, 0], [1, 1, 1])
t.fill_synthetic(20.0, 0)
proj_inv, views = bench.cameras(pkg, 0, 1)
t.set_proj_inv(proj_inv)
W, H, batch, iters = bench.W, bench.H, 4096, 20
rng = np.random.default_rng(1234)
vi_dev = pkg.DeviceBuffer(t, len(views) * 64); vi_dev.upload(np.stack(views).astype(np.float32))
picks = np.stack([rng.integers(0, len(views), (iters + 3) * batch), rng.integers(0, W * H, (iters + 3) * batch)], axis=1).astype(np.uint32)
picks_dev = pkg.DeviceBuffer(t, picks.nbytes); picks_dev.upload(picks)
ref = rng.random(((iters + 3) * batch, 4)).astype(np.float32)
ref_dev = pkg.DeviceBuffer(t, ref.nbytes); ref_dev.upload(ref)
o_dev, d_dev = pkg.DeviceBuffer(t, batch * 12), pkg.DeviceBuffer(t, batch * 12)
t.optimizerInit(); t.zeroGrad()


def gen(it):
    t._ck(t.lib.lrf_gen_rays_dev_async(t.h, batch, vi_dev.ptr, picks_dev.ptr + it * batch * 8, W, H, o_dev.ptr, d_dev.ptr))


def A(it):
    gen(it)
    t._ck(t.lib.lrf_rays_grad_dev_async(t.h, batch, o_dev.ptr, d_dev.ptr, ref_dev.ptr + it * batch * 16, 0))
    t._ck(t.lib.lrf_train_percell_async(t.h, it))


def B(it):
    gen(it)
    t._ck(t.lib.lrf_rays_grad_tv_dev_async(t.h, batch, o_dev.ptr, d_dev.ptr, ref_dev.ptr + it * batch * 16, 0))
    t._ck(t.lib.lrf_adam_zero_async(t.h, it))


for rep in range(2):
    for name, fn in (("A backward ; {TV -> Adam(d)} || Adam(c)", A), ("B backward || TV ; Adam + zero", B)):
        for it in range(3): fn(it)
        t.sync()
        t.timer_start()
        for it in range(3, 3 + iters): fn(it)
        print("%-44s %.4f ms per iteration" % (name, t.timer_stop_ms() / iters))
t.timer_start()
for it in range(10):
    t._ck(t.lib.lrf_rays_grad_tv_dev_async(t.h, batch, o_dev.ptr, d_dev.ptr, ref_dev.ptr, 0))
print("backward || TV alone                         %.4f ms" % (t.timer_stop_ms() / 10))
t.timer_start()
for it in range(10):
    t._ck(t.lib.lrf_rays_grad_dev_async(t.h, batch, o_dev.ptr, d_dev.ptr, ref_dev.ptr, 0))
print("backward alone                               %.4f ms" % (t.timer_stop_ms() / 10))
