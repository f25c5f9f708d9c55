"""torchrun --nproc-per-node N tools/mc_check.py [gs] — the multicast exchange step by step on real GPUs, synchronising after every
stage so that a failing stage names itself."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import bench

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = ge.package()
shard = bench.sharding()
gs = int(sys.argv[1]) if len(sys.argv) > 1 else 256
mode = sys.argv[2] if len(sys.argv) > 2 else "mc"
t = pkg.RayMarcherExample(local)
t.InitGrid(gs)
t.fill_synthetic(20.0, 0)
t.optimizerInit()
t.zeroGrad()
t.set_proj_inv(pkg.inverse4x4(pkg.perspective(45.0, 1.0, 0.1, 100.0)))
ex = shard.McExchange(t, dist, torch) if mode == "mc" else shard.PeerExchange(t, dist, torch)


def stage(name, fn, sync=True):
    try:
        fn()
        if sync:
            t.sync()
            torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            print("ok  ", name, flush=True)
    except Exception as e:
        print("FAIL rank %d %s: %s" % (rank, name, e), flush=True)
        os._exit(1)


n = 4096
rng = np.random.default_rng(rank)
views = np.stack([pkg.inverse4x4(pkg.orbit_view(k, 16)) for k in range(16)]).astype(np.float32)
vi_dev = pkg.DeviceBuffer(t, views.nbytes); vi_dev.upload(views)
picks = np.stack([rng.integers(0, 16, n), rng.integers(0, 800 * 800, n)], axis=1).astype(np.uint32)
pk = pkg.DeviceBuffer(t, picks.nbytes); pk.upload(picks)
rf = pkg.DeviceBuffer(t, n * 16); rf.upload(rng.random((n, 4)).astype(np.float32))
od, dd = pkg.DeviceBuffer(t, n * 12), pkg.DeviceBuffer(t, n * 12)
stage("gen rays", lambda: t._ck(t.lib.lrf_gen_rays_dev_async(t.h, n, vi_dev.ptr, pk.ptr, 800, 800, od.ptr, dd.ptr)))
stage("barrier", lambda: ex.barrier())
stage("adam exchange", lambda: t.adam_peer_async(0))
stage("zero", lambda: t.zero_grad_async())
stage("backward only", lambda: t._ck(t.lib.lrf_rays_grad_dev_async(t.h, n, od.ptr, dd.ptr, rf.ptr, 0)))
stage("tv slab", lambda: t.tv_grad_slab_async())
stage("percell", lambda: ex.percell(1, tv_done=True))
stage("backward || tv", lambda: t._ck(t.lib.lrf_rays_grad_tv_dev_async(t.h, n, od.ptr, dd.ptr, rf.ptr, 0)))
stage("percell", lambda: ex.percell(2, tv_done=True))
for k in range(3):
    stage("unsynchronised iterations x10 (%d)" % k, lambda: [(t._ck(t.lib.lrf_rays_grad_tv_dev_async(t.h, n, od.ptr, dd.ptr, rf.ptr, 0)), ex.percell(3 + i, tv_done=True)) for i in range(10)])
for k in range(3):
    stage("exchange alone x5 (%d)" % k, lambda: [ex.percell(40 + i, tv_done=True) for i in range(5)])
if mode == "mc":
    for unroll, bps in ((4, 0), (2, 0), (8, 0), (4, 8), (8, 4), (8, 8), (4, 16), (4, 0)):
        t.set_option("mc_unroll", unroll)
        t.set_option("mc_blocks_per_sm", bps)
        for i in range(3):
            ex.percell(100 + i, tv_done=True)
        t.sync(); dist.barrier()
        t.timer_start()
        for i in range(10):
            ex.percell(110 + i, tv_done=True)
        ms = torch.tensor([t.timer_stop_ms() / 10], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            print("exchange alone (barrier, multicast Adam pass, barrier, zero): unroll %d, blocks/SM cap %2d: %.3f ms" % (unroll, bps, float(ms.item())), flush=True)
    t.set_option("mc_unroll", 4); t.set_option("mc_blocks_per_sm", 0)
t.sync()
p = t.get_grid()
cs = float(np.abs(p.astype(np.float64)).sum())
alls = [None] * world
dist.all_gather_object(alls, cs)
if rank == 0:
    print("replica checksums", alls, "identical" if len(set(alls)) == 1 else "DIFFERENT")
ex.close()
t.close()
dist.destroy_process_group()
