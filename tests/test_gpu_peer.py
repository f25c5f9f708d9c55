"""Multi-GPU exchange over peer memory (lrf_train_percell_peer_async), exercised on ONE device with two contexts standing
in for two ranks (lrf_peer_attach_local; the IPC variant of the same kernel runs under torchrun, tools/peer_check.py).
Each "rank" back-propagates its shard of a ray batch; the slab owner sums the two gradients through the peer pointers,
applies TV + Adam and writes both replicas.  Expected values come from the oracle's Adam fed with the rank-ordered fp32
sum of the two gradients — bit-exact — and from a single-context full-batch run (same sum in another association)."""
import numpy as np
import pytest

import libs
from scenes import OracleScene, cameras

pytestmark = pytest.mark.gpu


def _mk(pkg, gs, grid):
    t = pkg.RayMarcherExample()
    t.InitGrid(gs)
    t.set_grid(grid)
    t.optimizerInit()
    t.zeroGrad()
    return t


@pytest.mark.parametrize("gs", [16, 32])
def test_peer_fused_percell_two_ranks(pkg, gs):
    grid = libs.lcg_grid(gs, 20.0)
    S = OracleScene(gs, grid)
    vi, pi = cameras(None, 2, 7)
    S.set_cam(vi, pi)
    o, d = S.gen_rays(24, 24)
    n = o.shape[0]
    ta, tb, full = _mk(pkg, gs, grid), _mk(pkg, gs, grid), _mk(pkg, gs, grid)
    ta.peer_attach_local(0, [ta, tb])
    tb.peer_attach_local(1, [ta, tb])
    S.O.orc_adam_init(S.h)
    for it in range(3):
        ref = np.random.default_rng(50 + it).random((n, 4)).astype(np.float32)
        half = n // 2
        ta.rays_grad(o[:half], d[:half], ref[:half])
        tb.rays_grad(o[half:], d[half:], ref[half:])
        ta.tv_grad_slab_async(); tb.tv_grad_slab_async()
        ta.sync(); tb.sync()
        ga, gb = ta.get_grad(), tb.get_grad()
        # TV went to the slab owner only: rank 0 owns z < gs/2 of the density plane
        zs = gs ** 3 // 2
        # full-batch reference run on one context: same rays, replicated TV
        full.rays_grad(o, d, ref)
        full.kernel2D_TVGrad()
        gf = full.get_grad().astype(np.float64)
        gsum = ga + gb                                            # fp32, rank order: what the kernel forms
        assert np.linalg.norm(gsum - gf) <= 1e-5 * np.linalg.norm(gf)
        # the exchange + Adam
        ta.adam_peer_async(it); tb.adam_peer_async(it)
        ta.sync(); tb.sync()
        pa, pb = ta.get_grid(), tb.get_grid()
        assert np.array_equal(pa.view(np.uint32), pb.view(np.uint32)), "replicas diverged"
        S.grad[:] = gsum
        S.O.orc_adam_step(S.h, it, 0)
        assert np.array_equal(pa.view(np.uint32), S.grid.view(np.uint32)), "peer Adam differs from the oracle's Adam on the summed gradient"
        # moments: rank r holds slab r
        Va, Sa = ta.get_adam_state()
        Vb, Sb = tb.get_adam_state()
        assert np.array_equal(Va[:zs], S.adam_v[:zs]) and np.array_equal(Vb[zs:], S.adam_v[zs:])
        assert np.array_equal(Sa[:zs], S.adam_s[:zs]) and np.array_equal(Sb[zs:], S.adam_s[zs:])
        full.optimizerStep(it)
        full.zeroGrad()
        assert np.allclose(full.get_grid(), pa, rtol=0, atol=0.31)   # Adam's first steps are ~ +-lr: same sign pattern up to cancellation noise
        ta.zero_grad_async(); tb.zero_grad_async()
        ta.sync(); tb.sync()
    for t in (ta, tb, full):
        t.close()
    S.close()


def test_peer_group_argument_checks(pkg):
    t = pkg.RayMarcherExample()
    t.InitGrid(16)
    with pytest.raises(pkg.LrfError):
        t.train_percell_peer_async(0)                # nothing attached
    with pytest.raises(pkg.LrfError):
        t.peer_attach_local(0, [t, t, t])            # 16 is not a multiple of 8 * 3
    t.peer_attach_local(0, [t])                      # world 1 degenerates to the local pass
    t.train_percell_peer_async(0)
    t.sync()
    t.close()


def test_tv_beside_backward_equals_sequential(pkg):
    """lrf_rays_grad_tv_dev_async (TV on a second stream, red.add) + lrf_adam_zero_async == backward; TV; Adam; zero"""
    gs = 32
    grid = libs.lcg_grid(gs, 20.0)
    S = OracleScene(gs, grid)
    vi, pi = cameras(None, 3, 7)
    S.set_cam(vi, pi)
    o, d = S.gen_rays(32, 32)
    n = o.shape[0]
    ref = np.random.default_rng(9).random((n, 4)).astype(np.float32)
    ta, tb = _mk(pkg, gs, grid), _mk(pkg, gs, grid)
    bufs = [pkg.DeviceBuffer(ta, a.nbytes) for a in (o, d, ref)]
    for b, a in zip(bufs, (o, d, ref)):
        b.upload(a)
    for it in range(2):
        ta.rays_grad_tv_dev_async(n, bufs[0].ptr, bufs[1].ptr, bufs[2].ptr)
        ta.sync()
        ga = ta.get_grad().astype(np.float64)
        tb.rays_grad(o, d, ref)
        tb.kernel2D_TVGrad()
        gb = tb.get_grad().astype(np.float64)
        assert np.linalg.norm(ga - gb) <= 1e-5 * np.linalg.norm(gb)
        tb.set_grad(ta.get_grad())            # same gradient bits into both optimisers
        ta.adam_zero_async(it)
        ta.sync()
        tb.optimizerStep(it)
        tb.zeroGrad()
        assert np.array_equal(ta.get_grid().view(np.uint32), tb.get_grid().view(np.uint32))
        assert not ta.get_grad().any()
    ta.close(); tb.close(); S.close()


def test_peer_flag_barrier_two_contexts(pkg):
    """the peer-memory barrier between two contexts on one device: both kernels must be in flight together (own streams)"""
    ta, tb = pkg.RayMarcherExample(), pkg.RayMarcherExample()
    ta.InitGrid(16); tb.InitGrid(16)
    ta.peer_attach_local(0, [ta, tb])
    tb.peer_attach_local(1, [ta, tb])
    for _ in range(20):
        ta.peer_barrier_async()
        tb.peer_barrier_async()
    assert not ta.peer_barrier_timed_out() and not tb.peer_barrier_timed_out()
    ta.close(); tb.close()


@pytest.mark.parametrize("gs", [16, 32])
def test_multicast_fused_percell(pkg, gs):
    """the exchange through NVLink multicast objects (lrf_mc_*, adam_mc_kernel: multimem.ld_reduce / multimem.st) with one context per
    visible device (at most 2; on a 1-GPU box the group has one member and the switch reduction degenerates to a read): replicas
    identical, parameters and moments equal the oracle's Adam fed with the summed gradient bit for bit (a two-term fp32 sum does not
    depend on the order), and the planes survive joining and leaving the group"""
    import torch
    world = min(2, torch.cuda.device_count())
    grid = libs.lcg_grid(gs, 20.0)
    S = OracleScene(gs, grid)
    vi, pi = cameras(None, 2, 7)
    S.set_cam(vi, pi)
    o, d = S.gen_rays(24, 24)
    n = o.shape[0]
    ts = []
    for r in range(world):
        t = pkg.RayMarcherExample(r)
        t.InitGrid(gs)
        t.set_grid(grid)
        t.optimizerInit()
        t.zeroGrad()
        ts.append(t)
    if not ts[0].mc_supported():
        pytest.skip("no NVLink multicast support on this device / driver")
    try:
        pkg.RayMarcherExample.mc_attach_local(ts)
    except pkg.LrfError as e:
        pytest.skip("multicast group could not be formed here: %s" % e)
    assert all(t.mc_attached() for t in ts)
    assert np.array_equal(ts[0].get_grid().view(np.uint32), np.asarray(grid, np.float32).reshape(-1, 13).view(np.uint32)), "planes changed while joining the group"
    S.O.orc_adam_init(S.h)
    for it in range(3):
        ref = np.random.default_rng(70 + it).random((n, 4)).astype(np.float32)
        cut = [n * r // world for r in range(world + 1)]
        for r, t in enumerate(ts):
            t.rays_grad(o[cut[r]:cut[r + 1]], d[cut[r]:cut[r + 1]], ref[cut[r]:cut[r + 1]])
            t.tv_grad_slab_async()
        for t in ts:
            t.sync()
        gsum = ts[0].get_grad()
        for t in ts[1:]:
            gsum = gsum + t.get_grad()
        for t in ts:
            t.adam_peer_async(it)
        for t in ts:
            t.sync()
        ps = [t.get_grid() for t in ts]
        for p in ps[1:]:
            assert np.array_equal(ps[0].view(np.uint32), p.view(np.uint32)), "replicas diverged"
        S.grad[:] = gsum
        S.O.orc_adam_step(S.h, it, 0)
        assert np.array_equal(ps[0].view(np.uint32), S.grid.view(np.uint32)), "multicast Adam differs from the oracle's Adam on the summed gradient"
        zs = gs ** 3 // world
        for r, t in enumerate(ts):
            Vr, Sr = t.get_adam_state()
            assert np.array_equal(Vr[zs * r:zs * (r + 1)], S.adam_v[zs * r:zs * (r + 1)])
            assert np.array_equal(Sr[zs * r:zs * (r + 1)], S.adam_s[zs * r:zs * (r + 1)])
        for t in ts:
            t.zero_grad_async()
        for t in ts:
            t.sync()
    last = ts[0].get_grid()
    for t in ts:
        t.peer_detach()
        assert not t.mc_attached()
        assert np.array_equal(t.get_grid().view(np.uint32), last.view(np.uint32)), "planes changed while leaving the group"
    # and the context keeps working on ordinary allocations
    ts[0].rays_grad(o, d, ref)
    ts[0].optimizerStep(3)
    for t in ts:
        t.close()
    S.close()


def test_peer_exchange_ipc_two_processes():
    """the CUDA-IPC path of the exchange under torchrun (2 ranks, one GPU each) against the NCCL all-reduce path; needs 2 GPUs"""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run by hand with gpurun --gpus 2; output kept in profiles/r1h_peer_check_n2.txt)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(root, "tools", "peer_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "peer replicas identical to rank 0: True" in r.stdout
