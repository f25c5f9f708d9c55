"""GPU parity for the larger BASELINE.json configurations:
  configs[3]  sparse box set (~10 % occupancy: density positive inside a sphere, negative outside, boxes kept by the
              reference's RebuildOctree criterion) — against the oracle at 64^3 / 128^3, and at the full 512^3 through
              size-independent properties (the streamlined march against the generic exact march, two independent
              implementations; empty rays; monotone hit counts);
  configs[4]  non-square frames (16:9) — against the oracle at a reduced frame size, same camera formula;
  the device-side synthetic generator against the host generator the oracle-side tests use."""
import numpy as np
import pytest

import libs
from scenes import OracleScene, cameras
from test_gpu_forward import _check_view

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("gs,D,mode", [(16, 20.0, 0), (32, 2.0, 0), (32, 20.0, 2), (32, 20.0, 1)])
def test_device_synthetic_grid_equals_host_generator(pkg, gs, D, mode):
    t = pkg.RayMarcherExample()
    t.InitGrid(gs)
    t.fill_synthetic(D, mode)
    host = libs.lcg_grid(gs, D, blob=(mode == 1), sphere=(mode == 2))
    dev = t.get_grid()
    assert np.array_equal(dev[:, 1:].view(np.uint32), host[:, 1:].view(np.uint32))
    if mode == 1:   # the blob factor goes through exp(): the host generator's is numpy's float32 exp, compare to 2 ulp
        assert np.allclose(dev[:, 0], host[:, 0], rtol=3e-7, atol=0)
    else:
        assert np.array_equal(dev[:, 0].view(np.uint32), host[:, 0].view(np.uint32))
    t.close()


@pytest.mark.parametrize("gs,wh", [(64, 48), (128, 40)])
def test_sparse_sphere_scene_against_oracle(pkg, gs, wh):
    """configs[3] scaled down: same generator, same box criterion, oracle side by side"""
    grid = libs.lcg_grid(gs, 20.0, sphere=True)
    S = OracleScene(gs, grid)
    S.O.orc_rebuild_boxes(S.h)
    t = pkg.RayMarcherExample()
    t.InitGrid(gs)
    t.fill_synthetic(20.0, 2)
    t.RebuildOctree()
    boxes = t.get_boxes()
    assert np.array_equal(boxes, S.boxes())
    occ = boxes.shape[0] / float((gs - 1) ** 3)
    assert 0.08 < occ < 0.12, occ
    for k in (0, 3):
        vi, pi = cameras(pkg, k, 7)
        S.set_cam(vi, pi)
        t.set_view_inv(vi)
        t.set_proj_inv(pi)
        nh, ns = _check_view(t, S, wh, wh, False, keep=64)
        assert nh.max() > 0 and (nh == 0).any()          # rays through the sphere and rays that miss every box
        assert (ns < nh).any()                           # early termination happens
    t.close()
    S.close()


def test_wide_frames_against_oracle(pkg):
    """configs[4] camera (perspective 45 deg, aspect 16:9) at 96x54 and a ragged 50x31 frame"""
    gs = 32
    grid = libs.lcg_grid(gs, 20.0)
    S = OracleScene(gs, grid)
    t = pkg.RayMarcherExample()
    t.InitGrid(gs)
    t.set_grid(grid)
    for (w, h, k) in ((96, 54, 1), (50, 31, 5)):
        vi, pi = cameras(pkg, k, 8, aspect=16.0 / 9.0)
        S.set_cam(vi, pi)
        t.set_view_inv(vi)
        t.set_proj_inv(pi)
        _check_view(t, S, w, h, False, keep=48)
    t.close()
    S.close()


def test_sparse512_properties(pkg):
    """configs[3] at full size: 512^3 nodes (7 GB per array, generated on the device), ~10 % of 511^3 boxes"""
    gs = 512
    t = pkg.RayMarcherExample()
    t.InitGrid(gs)
    t.fill_synthetic(20.0, 2)
    t.RebuildOctree()
    nb = t.box_count()
    occ = nb / float((gs - 1) ** 3)
    assert 0.095 < occ < 0.105, occ
    vi, pi = cameras(pkg, 2, 7)
    t.set_view_inv(vi)
    t.set_proj_inv(pi)
    w = h = 96
    nh_f, ns_f, rg_f, vox_f = t.render_debug(w, h, False, keep=32)
    t.set_option("force_generic_march", 1)
    nh_g, ns_g, rg_g, vox_g = t.render_debug(w, h, False, keep=32)
    t.set_option("force_generic_march", 0)
    assert np.array_equal(nh_f, nh_g) and np.array_equal(ns_f, ns_g) and np.array_equal(vox_f, vox_g)
    assert np.array_equal(rg_f[:, 3].view(np.uint32), rg_g[:, 3].view(np.uint32))
    assert np.abs(rg_f[:, :3] - rg_g[:, :3]).max() <= 1e-4
    # geometry of the scene: a ray crosses at most ~ sqrt(3) * 2r * gs boxes of the sphere, the frame's corner rays miss it
    assert nh_f.max() <= int(1.75 * 2 * 0.288 * gs) + 8 and nh_f.max() > 0.5 * 2 * 0.288 * gs
    img = np.zeros(w * h, np.uint32)
    t.kernel2D_RayMarch(img, w, h)
    assert ((img == 0) == (nh_f == 0)).all()             # ref :645,660: rays with an empty hit list give pixel 0
    assert (ns_f[nh_f > 0] >= 1).all() and (ns_f <= nh_f).all()
    t.close()


def test_grid_too_large_fails_cleanly(pkg):
    """the largest size the API accepts (1024^3 = 4 x 56 GB) does not fit in 180 GB: an error, not a half-built context"""
    t = pkg.RayMarcherExample()
    with pytest.raises(pkg.LrfError):
        t.InitGrid(1024)
    with pytest.raises(pkg.LrfError):
        t.kernel2D_RayMarch(np.zeros(16, np.uint32), 4, 4)     # no grid
    with pytest.raises(pkg.LrfError):
        t.InitGrid(48)                                          # not a power of two
    t.InitGrid(16)                                              # still usable
    img = np.zeros(64, np.uint32)
    vi, pi = cameras(pkg, 0, 7)
    t.set_view_inv(vi); t.set_proj_inv(pi)
    t.kernel2D_RayMarch(img, 8, 8)
    t.kernel2D_RayMarch(img, 0, 0)                              # empty frame: a no-op
    assert img.any()
    t.close()
