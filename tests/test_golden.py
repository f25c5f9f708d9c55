"""Replays the committed golden vectors (tests/golden/*.npz — outputs of the UNMODIFIED reference, produced by
tests/gen_golden.py in the dev container) against
  * the CPU oracle           (not gpu: runs everywhere; pins the oracle without /root/reference), and
  * the CUDA path via C-ABI  (gpu: the parity test proper; /root/reference does not exist on the GPU box).
Bars: hit counts, shaded counts, hit ids/voxels bit-exact; oracle floats bit-exact; CUDA RGB within 1e-4 absolute,
transmittance bit-exact, packed pixels within 1 LSB on < 0.1 % of channels.
"""
import os

import numpy as np
import pytest

import libs
from libs import fp, u32p, i64p
from scenes import OracleScene

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FWD = ["fwd_gs16_opaque", "fwd_gs16_thin", "fwd_gs16_grey", "fwd_gs32_opaque", "fwd_gs64_blob"]
RGB_TOL = 1e-4


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def dense_xyz_of_ids(ids, gs):
    n = gs - 1
    xyz = np.stack([ids % n, (ids // n) % n, ids // (n * n)], axis=-1).astype(np.int32)
    return np.where(ids[..., None] >= 0, xyz, -1)


# ------------------------------------------------------------------------------------------------ CPU: oracle
@pytest.mark.parametrize("name", FWD)
def test_oracle_matches_golden_forward(orc, name):
    G = load(name)
    gs, wh, grey, keep = int(G["gs"]), int(G["wh"]), int(G["grey"]), int(G["keep"])
    S = OracleScene(gs, libs.lcg_grid(gs, float(G["D"]), blob=bool(G["blob"])))
    for k in G["views"]:
        S.set_cam(G["view_inv_%d" % k], G["proj_inv_%d" % k])
        n = wh * wh
        nh, ns = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        rg, ids = np.zeros((n, 5), np.float32), np.zeros((n, keep), np.int64)
        orc.orc_render_debug(S.h, wh, wh, grey, u32p(nh), u32p(ns), fp(rg), i64p(ids), keep)
        assert np.array_equal(nh, G["n_hit_%d" % k]) and np.array_equal(ns, G["n_shaded_%d" % k])
        assert np.array_equal(ids, G["ids_%d" % k])
        assert np.array_equal(bits(rg), bits(G["rgbtc_%d" % k]))
        assert np.array_equal(S.render(wh, wh, grey), G["pixels_%d" % k])
    S.close()


def test_oracle_matches_golden_sparse(orc):
    G = load("sparse_upsampled")
    gs = int(G["gs0"])
    grid = libs.lcg_grid(gs, 20.0, blob=True)
    grid[:, 0] -= 0.4
    S = OracleScene(gs, grid)
    for level in range(2):
        orc.orc_upsample_grid(S.h)
        orc.orc_upsample_boxes(S.h)
        assert np.array_equal(S.boxes(), G["boxes_%d" % level])
        g = S.grid
        assert np.array_equal(bits(g[:: max(1, g.shape[0] // 512)]), bits(G["grid_sample_%d" % level]))
        assert g.astype(np.float64).sum() == float(G["grid_sum_%d" % level][0])
        S.set_cam(G["view_inv_%d" % level], G["proj_inv_%d" % level])
        nh, ns = np.zeros(1600, np.uint32), np.zeros(1600, np.uint32)
        rg, ids = np.zeros((1600, 5), np.float32), np.zeros((1600, 40), np.int64)
        orc.orc_render_debug(S.h, 40, 40, 0, u32p(nh), u32p(ns), fp(rg), i64p(ids), 40)
        assert np.array_equal(nh, G["n_hit_%d" % level]) and np.array_equal(ns, G["n_shaded_%d" % level])
        assert np.array_equal(ids, G["ids_%d" % level]) and np.array_equal(bits(rg), bits(G["rgbtc_%d" % level]))
    S.close()


def _adam_grads(which, gs):
    rng = np.random.default_rng(100 + which)
    return [(rng.standard_normal((gs ** 3, 13)) * (rng.random((gs ** 3, 13)) < 0.4)).astype(np.float32) for _ in range(3)]


def test_oracle_matches_golden_percell(orc):
    G = load("percell")
    gs = int(G["gs"])
    grid = libs.lcg_grid(gs, 20.0)
    for which in (0, 1, 2):
        S = OracleScene(gs, grid)
        for it, g in enumerate(_adam_grads(which, gs)):
            S.grad[:] = g
            orc.orc_adam_step(S.h, it, which)
        assert np.array_equal(bits(S.grid), bits(G["adam_p_%d" % which]))
        assert np.array_equal(bits(S.adam_v), bits(G["adam_v_%d" % which])) and np.array_equal(bits(S.adam_s), bits(G["adam_s_%d" % which]))
        S.close()
    S = OracleScene(gs, grid)
    tv = G["tv_O0"]
    for z in range(gs - 1):
        for y in range(gs - 1):
            for x in range(gs - 1):
                assert np.float32(orc.orc_tv(S.h, x, y, z)).view(np.uint32) == tv[z, y, x].view(np.uint32)
    S.close()


# ------------------------------------------------------------------------------------------------ GPU: CUDA path
@pytest.mark.gpu
@pytest.mark.parametrize("name", FWD)
def test_cuda_matches_golden_forward(pkg, name):
    G = load(name)
    gs, wh, grey, keep = int(G["gs"]), int(G["wh"]), int(G["grey"]), int(G["keep"])
    t = pkg.RayMarcherExample()
    t.InitGrid(gs)
    t.SetBoundingBox([0, 0, 0], [1, 1, 1])
    t.set_grid(libs.lcg_grid(gs, float(G["D"]), blob=bool(G["blob"])))
    for k in G["views"]:
        t.set_view_inv(G["view_inv_%d" % k])
        t.set_proj_inv(G["proj_inv_%d" % k])
        nh, ns, rg, vox = t.render_debug(wh, wh, bool(grey), keep)
        assert np.array_equal(nh, G["n_hit_%d" % k]), "hit counts"
        assert np.array_equal(ns, G["n_shaded_%d" % k]), "shaded-sample counts"
        assert np.array_equal(vox, dense_xyz_of_ids(G["ids_%d" % k].astype(np.int64), gs)), "voxel indices / order"
        ref = G["rgbtc_%d" % k]
        assert np.abs(ref[:, :3] - rg[:, :3]).max() <= RGB_TOL
        assert np.array_equal(bits(ref[:, 3]), bits(rg[:, 3])), "transmittance"
        img = np.zeros(wh * wh, np.uint32)
        t.kernel2D_RayMarch(img, wh, wh, bool(grey))
        ch_r = np.stack([(G["pixels_%d" % k] >> s) & 255 for s in (0, 8, 16, 24)]).astype(np.int32)
        ch_g = np.stack([(img >> s) & 255 for s in (0, 8, 16, 24)]).astype(np.int32)
        assert np.abs(ch_r - ch_g).max() <= 1 and (ch_r != ch_g).mean() < 1e-3
    t.close()


@pytest.mark.gpu
def test_cuda_matches_golden_sparse(pkg):
    G = load("sparse_upsampled")
    gs = int(G["gs0"])
    grid = libs.lcg_grid(gs, 20.0, blob=True)
    grid[:, 0] -= 0.4
    t = pkg.RayMarcherExample()
    t.InitGrid(gs)
    t.set_grid(grid)
    for level in range(2):
        t.UpsampleGrid()
        t.InitGrad()
        t.optimizerInit()
        t.UpsampleOctree()
        bx = G["boxes_%d" % level].astype(np.int32)
        assert np.array_equal(t.get_boxes(), bx)
        g = t.get_grid()
        assert np.array_equal(bits(g[:: max(1, g.shape[0] // 512)]), bits(G["grid_sample_%d" % level]))
        t.set_view_inv(G["view_inv_%d" % level])
        t.set_proj_inv(G["proj_inv_%d" % level])
        nh, ns, rg, vox = t.render_debug(40, 40, False, 40)
        assert np.array_equal(nh, G["n_hit_%d" % level]) and np.array_equal(ns, G["n_shaded_%d" % level])
        ids = G["ids_%d" % level].astype(np.int64)
        assert np.array_equal(vox, np.where(ids[..., None] >= 0, bx[np.maximum(ids, 0)], -1))
        assert np.abs(G["rgbtc_%d" % level][:, :3] - rg[:, :3]).max() <= RGB_TOL
    t.close()


@pytest.mark.gpu
def test_cuda_matches_golden_adam(pkg):
    G = load("percell")
    gs = int(G["gs"])
    grid = libs.lcg_grid(gs, 20.0)
    for which in (0, 1, 2):
        t = pkg.RayMarcherExample()
        t.InitGrid(gs)
        t.set_grid(grid)
        for it, g in enumerate(_adam_grads(which, gs)):
            t.set_grad(g)
            {0: t.optimizerStep, 1: t.optimizerStepDensity, 2: t.optimizerStepSH}[which](it)
        V, S = t.get_adam_state()
        assert np.array_equal(bits(t.get_grid()), bits(G["adam_p_%d" % which]))
        assert np.array_equal(bits(V), bits(G["adam_v_%d" % which])) and np.array_equal(bits(S), bits(G["adam_s_%d" % which]))
        t.close()
