"""N > 1 host logic on CPU: world_size-2 `gloo` process group.  Each rank back-propagates ITS shard of a packed ray
batch with the CPU oracle, the gradient planes are summed with the same all-reduce call the GPU harness issues over
NCCL, and the result must equal the single-process full-batch gradient (data parallelism == a larger batch, because
the reference's loss is an un-normalised sum over rays); view sharding must deal every view exactly once."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "1"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import libs
    from libs import fp
    from conftest import load_package
    from scenes import OracleScene, cameras
    shard = __import__("importlib").import_module("importlib").util.spec_from_file_location(
        "lrf_sharding", os.path.join(ROOT, "literf-radiancefields_b200", "sharding.py"))
    mod = __import__("importlib").util.module_from_spec(shard)
    shard.loader.exec_module(mod)

    gs = 16
    S = OracleScene(gs, libs.lcg_grid(gs, 20.0))
    vi, pi = cameras(None, 2, 7)
    S.set_cam(vi, pi)
    o, d = S.gen_rays(24, 24)
    ref = np.random.default_rng(5).random((o.shape[0], 4)).astype(np.float32)
    b, e = mod.shard_rays(o.shape[0], rank, world)
    S.O.orc_zero_grad(S.h)
    S.O.orc_rays_grad(S.h, e - b, fp(np.ascontiguousarray(o[b:e])), fp(np.ascontiguousarray(d[b:e])), fp(np.ascontiguousarray(ref[b:e])), 0)
    g = torch.from_numpy(S.grad.copy())
    mod.allreduce_gradient(g, dist)
    # replicated per-cell pass on the summed gradient keeps the replicas identical
    S.grad[:] = g.numpy()
    S.O.orc_tv_grad(S.h)
    S.O.orc_adam_step(S.h, 0, 0)
    np.save(os.path.join(out_dir, "grad_%d.npy" % rank), g.numpy())
    np.save(os.path.join(out_dir, "grid_%d.npy" % rank), S.grid.copy())
    views = torch.zeros(100, dtype=torch.int32)
    for v in mod.shard_views(100, rank, world):
        views[v] += 1
    dist.all_reduce(views)
    assert bool((views == 1).all())
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_gradient_equals_full_batch(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, HERE)
    import libs
    from libs import fp
    from scenes import OracleScene, cameras
    gs = 16
    S = OracleScene(gs, libs.lcg_grid(gs, 20.0))
    vi, pi = cameras(None, 2, 7)
    S.set_cam(vi, pi)
    o, d = S.gen_rays(24, 24)
    ref = np.random.default_rng(5).random((o.shape[0], 4)).astype(np.float32)
    S.O.orc_zero_grad(S.h)
    S.O.orc_rays_grad(S.h, o.shape[0], fp(o), fp(d), fp(ref), 0)
    full = S.grad.astype(np.float64)
    g0, g1 = np.load(tmp_path / "grad_0.npy"), np.load(tmp_path / "grad_1.npy")
    assert np.array_equal(g0, g1)                                   # all-reduce leaves identical replicas
    assert np.abs(full).max() > 0
    assert np.linalg.norm(g0 - full) <= 1e-5 * np.linalg.norm(full)   # same sum, different association
    assert np.array_equal(np.load(tmp_path / "grid_0.npy"), np.load(tmp_path / "grid_1.npy"))
    S.close()


def test_shard_rays_partition():
    spec = __import__("importlib").util.spec_from_file_location("lrf_sharding", os.path.join(ROOT, "literf-radiancefields_b200", "sharding.py"))
    mod = __import__("importlib").util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for n in (0, 1, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            cuts = [mod.shard_rays(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            assert max(e - b for b, e in cuts) - min(e - b for b, e in cuts) <= 1


def _band_worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "1"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib.util
    import libs
    from scenes import OracleScene, cameras
    spec = importlib.util.spec_from_file_location("lrf_sharding", os.path.join(ROOT, "literf-radiancefields_b200", "sharding.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    gs, w, h = 16, 40, 30                     # 30 rows: bands of 4-row units, the last one ragged
    S = OracleScene(gs, libs.lcg_grid(gs, 20.0))
    vi, pi = cameras(None, 3, 7)
    S.set_cam(vi, pi)
    full = S.render(w, h, False).astype(np.int64)
    y0, y1 = mod.shard_rows(h, rank, world)
    mine = np.full(w * h, -1, np.int64)       # this rank only "renders" its band
    mine[y0 * w:y1 * w] = full[y0 * w:y1 * w]
    frame = torch.from_numpy(mine)
    mod.gather_frame(frame, h, w, dist)
    if rank == 0:
        np.save(os.path.join(out_dir, "frame.npy"), frame.numpy())
        np.save(os.path.join(out_dir, "full.npy"), full)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_row_band_frame_gather(tmp_path, world):
    """latency mode (one frame over several GPUs): bands tile the rows exactly once, rank 0 assembles the reference's frame"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("lrf_sharding", os.path.join(ROOT, "literf-radiancefields_b200", "sharding.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for h in (1, 3, 4, 30, 800, 1080, 2160):
        for wd in (1, 2, 3, 4, 8):
            cover = np.zeros(h, np.int32)
            for r in range(wd):
                y0, y1 = mod.shard_rows(h, r, wd)
                assert 0 <= y0 <= y1 <= h and (y0 % 4 == 0 or y0 == h)
                cover[y0:y1] += 1
            assert (cover == 1).all()
    mp.spawn(_band_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert np.array_equal(np.load(os.path.join(str(tmp_path), "frame.npy")), np.load(os.path.join(str(tmp_path), "full.npy")))
