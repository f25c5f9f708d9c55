"""GPU parity of the register-window forward march (lrf_lean.cuh; ref kernel2D_RayMarch example_tracer.cpp:627-663, the sample loop
RayGridIntersection :121-194) — the kernel `kernel2D_RayMarch` runs for dense colour grids.

Bar: per pixel the shaded-sample count before the reference's `throughput < 0.1` break (:131) and the final transmittance equal the
ORACLE's bit for bit; every packed pixel equals the oracle's up to the 1e-4 colour tolerance (<= 1 LSB after the reference's
(uint)(c*255) truncation); and the older gather kernel (option lean_render = 0) still produces the same frame."""
import numpy as np
import pytest

import libs
from scenes import OracleScene, cameras

pytestmark = pytest.mark.gpu

FALLBACK = 0xFFFFFFFF


def _channels(img):
    return np.stack([(img >> s) & 255 for s in (0, 8, 16, 24)]).astype(np.int32)


def _check(t, S, w, h):
    nh_o, ns_o, rg_o, _ = S.render_debug(w, h, False, 0)
    img_o = S.render(w, h, False)
    img_g, ns_g, thr_g = t.render_lean_debug(w, h)
    dec = (ns_g != FALLBACK) & (nh_o > 0)
    assert np.array_equal(ns_g[dec], ns_o[dec]), "shaded counts differ on %d rays" % (ns_g[dec] != ns_o[dec]).sum()
    assert np.array_equal(thr_g[dec].view(np.uint32), rg_o[dec, 3].view(np.uint32)), "transmittance not bit-exact"
    assert (ns_g[(ns_g != FALLBACK) & (nh_o == 0)] == 0).all()
    ch_o, ch_g = _channels(img_o), _channels(img_g)
    assert np.abs(ch_o - ch_g).max() <= 1
    assert np.array_equal(ch_o[3], ch_g[3])           # alpha: 255 on rays with a hit, 0 otherwise
    assert (ch_o != ch_g).mean() < 1e-3
    # the plain call renders the same frame; so does the older gather kernel, up to the colour tolerance
    plain = np.zeros(w * h, np.uint32)
    t.kernel2D_RayMarch(plain, w, h)
    assert np.array_equal(plain, img_g)
    t.set_option("lean_render", 0)
    old = np.zeros(w * h, np.uint32)
    t.kernel2D_RayMarch(old, w, h)
    t.set_option("lean_render", 1)
    assert np.abs(_channels(old) - ch_g).max() <= 1
    return int((ns_g == FALLBACK).sum()), int(ns_o.sum())


@pytest.mark.parametrize("gs,D,wh", [(4, 6.0, 48), (8, 20.0, 48), (16, 20.0, 64), (32, 2.0, 96), (64, 20.0, 128), (128, 20.0, 160)])
def test_lean_orbit_views(pkg, gs, D, wh):
    grid = libs.lcg_grid(gs, D)
    S = OracleScene(gs, grid)
    t = pkg.RayMarcherExample()
    t.InitGrid(gs)
    t.SetBoundingBox([0, 0, 0], [1, 1, 1])
    t.set_grid(grid)
    shaded = 0
    for k in (0, 2, 5) if gs >= 128 else range(7):
        vi, pi = cameras(pkg, k, 7)
        S.set_cam(vi, pi)
        t.set_view_inv(vi)
        t.set_proj_inv(pi)
        shaded += _check(t, S, wh, wh)[1]
    assert shaded > 0
    t.close()
    S.close()


def test_lean_symmetric_camera_ties(pkg):
    """a camera on the grid's axis of symmetry: the rays of the image diagonals have |dx| == |dy| and cross x and y planes at exactly the
    same distance at every step (3 boxes per event, ordered by id: the octree's tie order, ref external/Octree/octree.h:856-860), the centre
    row / column have a zero direction component (generic fallback)"""
    gs, w = 32, 64
    grid = libs.lcg_grid(gs, 6.0)
    S = OracleScene(gs, grid)
    t = pkg.RayMarcherExample()
    t.InitGrid(gs)
    t.SetBoundingBox([0, 0, 0], [1, 1, 1])
    t.set_grid(grid)
    _, pi = cameras(pkg, 0, 7)
    ties = 0
    for eye_z, flip in ((-1.5, 1.0), (2.5, -1.0)):
        # view_inv = camera-to-world: camera at (0.5, 0.5, eye_z) looking along +-z, no rotation about the axis (column-major like LiteMath)
        vi = np.zeros(16, np.float32)
        vi[0] = 1.0; vi[5] = 1.0; vi[10] = -flip; vi[15] = 1.0
        vi[12], vi[13], vi[14] = 0.5, 0.5, eye_z
        S.set_cam(vi, pi)
        t.set_view_inv(vi)
        t.set_proj_inv(pi)
        nh_o, ns_o, _, _ = S.render_debug(w, w, False, 0)
        n_fb, _ = _check(t, S, w, w)
        # the diagonal rays really do see ties: more hits than a ray crossing one face at a time could collect
        diag = np.arange(w) * w + np.arange(w)
        ties += int((nh_o[diag] > 0).sum())
        assert n_fb <= 2 * w
    assert ties > 0
    t.close()
    S.close()


def _sym_view(eye_z, flip):
    vi = np.zeros(16, np.float32)
    vi[0] = 1.0; vi[5] = 1.0; vi[10] = -flip; vi[15] = 1.0
    vi[12], vi[13], vi[14] = 0.5, 0.5, eye_z
    return vi


@pytest.mark.parametrize("mode", ["dense", "grey", "sparse", "staged"])
def test_render_views_equals_view_by_view(pkg, mode):
    """lrf_render_views_dev_async: the frames of one batch call are the frames kernel2D_RayMarch renders view by view, bit for bit —
    one launch over all views (the symmetric camera's centre row / column travel through the batch's fallback list), the register-window
    kernel for dense colour frames, the gather kernel for greyscale and sparse sets; the context's own view matrix is left alone"""
    gs, w, h = 32, 72, 40          # not a multiple of the 16 x 8 block: partial tiles in every view
    grid = libs.lcg_grid(gs, 6.0)
    t = pkg.RayMarcherExample()
    t.InitGrid(gs)
    t.SetBoundingBox([0, 0, 0], [1, 1, 1])
    t.set_grid(grid)
    if mode == "sparse":
        t.RebuildOctree()
    if mode == "staged":
        t.set_option("staged_render", 1)            # the opt-in TMA-staged kernel: the batch call runs it view by view
    grey = mode == "grey"
    views = [cameras(pkg, k, 7)[0] for k in (0, 3, 5)] + [_sym_view(-1.5, 1.0)] + [cameras(pkg, 6, 7)[0]]
    _, pi = cameras(pkg, 0, 7)
    t.set_proj_inv(pi)
    t.set_view_inv(views[1])
    frames = pkg.DeviceBuffer(t, len(views) * w * h * 4)
    l0 = t.launch_count()
    t.render_views_dev_async(views, frames.ptr, w, h, grey)
    launches = t.launch_count() - l0
    got = np.zeros(len(views) * w * h, np.uint32)
    frames.download(got)
    got = got.reshape(len(views), w * h)
    if mode == "staged":
        assert launches == 2 * len(views) + 1       # two kernels per view + the record bake
    else:
        assert launches == 2                        # the forward kernel over all views + the fallback kernel
    one = np.zeros(w * h, np.uint32)
    t.kernel2D_RayMarch(one, w, h, grey)            # the context's view is still views[1]
    assert np.array_equal(one, got[1])
    for k, vi in enumerate(views):
        t.set_view_inv(vi)
        t.kernel2D_RayMarch(one, w, h, grey)
        assert np.array_equal(one, got[k]), "view %d differs on %d pixels" % (k, (one != got[k]).sum())
        assert one.any()
    t.close()
