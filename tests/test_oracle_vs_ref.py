"""Pins the CPU oracle (oracle/lrf_oracle.cpp) against the UNMODIFIED reference compiled from
/root/reference (oracle/_ref/libref_{O3,O0}.so, built by oracle/Makefile through ref_harness.cpp).

Runs only where the reference build exists (the dev container; oracle/_ref/*.so also travels to the GPU
box with the snapshot, so the same tests run there).  Everything on the forward path is compared
BIT-FOR-BIT: camera matrices, rays, hit lists, shading outputs, packed pixels, TV value, Adam,
UpsampleGrid / UpsampleOctree / RebuildOctree.  The backward cannot be compiled (Enzyme) — see
test_oracle_backward_fd.py for how it is certified instead.
"""
import numpy as np
import pytest

import libs
from libs import fp, dp, u32p, u64p, i64p, np_view
from scenes import OracleScene

pytestmark = pytest.mark.skipif(not libs.have_ref("O3"), reason="oracle/_ref not built (needs /root/reference)")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


class RefScene:
    def __init__(self, gs, grid=None, kind="O3", boxes=True):
        self.R = libs.ref(kind)
        self.h = self.R.ref_create()
        (self.R.ref_init_grid if boxes else self.R.ref_init_grid_nobox)(self.h, gs)
        if grid is not None:
            self.grid[:] = np.asarray(grid, np.float32).reshape(-1, 13)

    def _n(self):
        return int(self.R.ref_grid_size(self.h)) ** 3 * 13

    @property
    def grid(self):
        return np_view(self.R.ref_grid_ptr(self.h), self._n()).reshape(-1, 13)

    @property
    def grad(self):
        return np_view(self.R.ref_grad_ptr(self.h), self._n()).reshape(-1, 13)

    @property
    def adam_v(self):
        return np_view(self.R.ref_adam_v_ptr(self.h), self._n()).reshape(-1, 13)

    @property
    def adam_s(self):
        return np_view(self.R.ref_adam_s_ptr(self.h), self._n()).reshape(-1, 13)

    def set_cam(self, k, n_views, fov=45.0, aspect=1.0):
        proj, view = np.zeros(16, np.float32), np.zeros(16, np.float32)
        self.R.ref_make_perspective(fov, aspect, 0.1, 100.0, fp(proj))
        self.R.ref_make_orbit_view(k, n_views, fp(view))
        self.R.ref_set_proj_matrix(self.h, fp(proj))
        self.R.ref_set_view_matrix(self.h, fp(view))
        vi, pi = np.zeros(16, np.float32), np.zeros(16, np.float32)
        self.R.ref_get_inv_matrices(self.h, fp(vi), fp(pi))
        return vi, pi

    def boxes_xyz(self):
        n = self.R.ref_box_count(self.h)
        b = np.zeros((n, 6), np.float64)
        self.R.ref_get_boxes(self.h, dp(b))
        gs = int(self.R.ref_grid_size(self.h))
        xyz = np.rint(b[:, :3] * gs).astype(np.int32)
        # the boxes really are unit voxels of the lattice with float-rounded planes
        assert np.array_equal((xyz.astype(np.float32) / np.float32(gs)).astype(np.float64), b[:, :3])
        assert np.array_equal(((xyz + 1).astype(np.float32) / np.float32(gs)).astype(np.float64), b[:, 3:])
        return xyz

    def render_debug(self, w, h, grey=False, keep=0):
        n = w * h
        nh, ns = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        rg = np.zeros((n, 5), np.float32)
        ids = np.zeros((n, max(keep, 1)), np.int64)
        self.R.ref_render_debug(self.h, w, h, int(grey), u32p(nh), u32p(ns), fp(rg), i64p(ids), keep)
        return nh, ns, rg, ids

    def close(self):
        self.R.ref_destroy(self.h)


def test_camera_helpers_bit_exact(orc):
    R = libs.ref("O3")
    for (fov, asp) in [(45.0, 1.0), (45.0, 16.0 / 9.0), (60.0, 0.75)]:
        a, b = np.zeros(16, np.float32), np.zeros(16, np.float32)
        R.ref_make_perspective(fov, asp, 0.1, 100.0, fp(a))
        orc.orc_perspective(fov, asp, 0.1, 100.0, fp(b))
        assert np.array_equal(bits(a), bits(b))
        ia, ib = np.zeros(16, np.float32), np.zeros(16, np.float32)
        R.ref_inverse4x4(fp(a), fp(ia))
        orc.orc_inverse4x4(fp(b), fp(ib))
        assert np.array_equal(bits(ia), bits(ib))
    for n_views in (7, 100, 800):
        for k in range(0, n_views, max(1, n_views // 13)):
            a, b = np.zeros(16, np.float32), np.zeros(16, np.float32)
            R.ref_make_orbit_view(k, n_views, fp(a))
            orc.orc_orbit_view(k, n_views, fp(b))
            assert np.array_equal(bits(a), bits(b)), (k, n_views)
            ia, ib = np.zeros(16, np.float32), np.zeros(16, np.float32)
            R.ref_inverse4x4(fp(a), fp(ia))
            orc.orc_inverse4x4(fp(b), fp(ib))
            assert np.array_equal(bits(ia), bits(ib))


def test_init_grid_boxes_and_hparams(orc):
    for gs in (8, 16):
        Rf, S = RefScene(gs), OracleScene(gs)
        assert np.array_equal(Rf.grid, S.grid)
        assert np.array_equal(Rf.boxes_xyz(), S.boxes())          # id = x + y(gs-1) + z(gs-1)^2, example_tracer.h:77
        hp = np.zeros(4, np.float32)
        Rf.R.ref_get_adam_hparams(Rf.h, fp(hp))
        assert np.array_equal(hp, np.array([0.15, 0.9, 0.999, 1e-8], np.float32))
        Rf.close()
        S.close()


@pytest.mark.parametrize("gs,D,wh,views", [(16, 20.0, 48, (0, 1, 3, 6)), (16, 2.0, 48, (2, 5)), (32, 20.0, 40, (0, 4))])
def test_dense_forward_bit_exact(orc, gs, D, wh, views):
    grid = libs.lcg_grid(gs, D)
    Rf, S = RefScene(gs, grid), OracleScene(gs, grid)
    keep = 3 * gs
    for k in views:
        vi, pi = Rf.set_cam(k, 7)
        S.set_cam(vi, pi)
        # rays
        o_r, d_r, o_o, d_o = (np.zeros(3, np.float32) for _ in range(4))
        for (x, y) in [(0, 0), (wh - 1, 0), (wh // 2, wh // 3), (5, wh - 1)]:
            Rf.R.ref_gen_ray(Rf.h, x, y, wh, wh, fp(o_r), fp(d_r))
            orc.orc_gen_ray(S.h, x, y, wh, wh, fp(o_o), fp(d_o))
            assert np.array_equal(bits(o_r), bits(o_o)) and np.array_equal(bits(d_r), bits(d_o))
        nh_r, ns_r, rg_r, ids_r = Rf.render_debug(wh, wh, False, keep)
        n = wh * wh
        nh_o, ns_o = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        rg_o = np.zeros((n, 5), np.float32)
        ids_o = np.zeros((n, keep), np.int64)
        orc.orc_render_debug(S.h, wh, wh, 0, u32p(nh_o), u32p(ns_o), fp(rg_o), i64p(ids_o), keep)
        assert np.array_equal(nh_r, nh_o), "hit counts"
        assert np.array_equal(ids_r, ids_o), "hit order / ids"
        assert np.array_equal(ns_r, ns_o), "shaded-sample counts (throughput break)"
        assert np.array_equal(bits(rg_r), bits(rg_o)), "rgb / T / cauchy bit-exact"
        assert nh_r.max() > gs and (ns_r < nh_r).any() == (D > 10)
        img_r, img_o = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        Rf.R.ref_render(Rf.h, u32p(img_r), wh, wh, 0)
        orc.orc_render(S.h, u32p(img_o), wh, wh, 0)
        assert np.array_equal(img_r, img_o), "packed pixels"
    Rf.close()
    S.close()


def test_greyscale_and_ragged_image(orc):
    gs = 16
    grid = libs.lcg_grid(gs, 20.0)
    Rf, S = RefScene(gs, grid), OracleScene(gs, grid)
    vi, pi = Rf.set_cam(3, 7)
    S.set_cam(vi, pi)
    for (w, h, grey) in [(33, 17, False), (24, 24, True), (1, 1, False)]:
        a, b = np.zeros(w * h, np.uint32), np.zeros(w * h, np.uint32)
        Rf.R.ref_render(Rf.h, u32p(a), w, h, int(grey))
        orc.orc_render(S.h, u32p(b), w, h, int(grey))
        assert np.array_equal(a, b)
    Rf.close()
    S.close()


def test_sparse_set_produced_the_reference_way(orc):
    """8^3 blob -> UpsampleGrid + UpsampleOctree twice (main.cpp:125-133): grids, surviving boxes, their id
    order (parent-major, child order of example_tracer.cpp:363-441) and renders over the sparse set (the
    `delta > sqrt(3) step` rule at :145-146 is live here) are bit-exact; then RebuildOctree."""
    gs = 8
    grid = libs.lcg_grid(gs, 20.0, blob=True)
    grid[:, 0] -= 0.4
    Rf, S = RefScene(gs, grid), OracleScene(gs, grid)
    for level in range(2):
        Rf.R.ref_upsample_grid(Rf.h)
        Rf.R.ref_upsample_octree(Rf.h)
        orc.orc_upsample_grid(S.h)
        orc.orc_upsample_boxes(S.h)
        assert np.array_equal(bits(Rf.grid), bits(S.grid))
        bx = Rf.boxes_xyz()
        assert np.array_equal(bx, S.boxes())
        assert 0 < bx.shape[0] < (orc.orc_grid_size(S.h) - 1) ** 3
        vi, pi = Rf.set_cam(1 + level, 7)
        S.set_cam(vi, pi)
        w = 40
        nh_r, ns_r, rg_r, ids_r = Rf.render_debug(w, w, False, 40)
        nh_o, ns_o = np.zeros(w * w, np.uint32), np.zeros(w * w, np.uint32)
        rg_o, ids_o = np.zeros((w * w, 5), np.float32), np.zeros((w * w, 40), np.int64)
        orc.orc_render_debug(S.h, w, w, 0, u32p(nh_o), u32p(ns_o), fp(rg_o), i64p(ids_o), 40)
        assert np.array_equal(nh_r, nh_o) and np.array_equal(ids_r, ids_o) and np.array_equal(ns_r, ns_o)
        assert np.array_equal(bits(rg_r), bits(rg_o))
    g2 = Rf.grid.copy()
    g2[:, 0] -= 0.05
    Rf.grid[:] = g2
    S.grid[:] = g2
    Rf.R.ref_rebuild_octree(Rf.h)
    orc.orc_rebuild_boxes(S.h)
    assert np.array_equal(Rf.boxes_xyz(), S.boxes())
    Rf.close()
    S.close()


def test_ray_loss_value(orc):
    """RayGridLoss (example_tracer.cpp:258-268), the function Enzyme differentiates, on explicit id lists"""
    gs = 16
    grid = libs.lcg_grid(gs, 20.0)
    Rf, S = RefScene(gs, grid), OracleScene(gs, grid)
    vi, pi = Rf.set_cam(2, 7)
    S.set_cam(vi, pi)
    rng = np.random.default_rng(7)
    o, d = np.zeros(3, np.float32), np.zeros(3, np.float32)
    ids = np.zeros(256, np.uint64)
    checked = 0
    for _ in range(200):
        x, y = rng.integers(0, 64, 2)
        orc.orc_gen_ray(S.h, int(x), int(y), 64, 64, fp(o), fp(d))
        n = orc.orc_ray_hits(S.h, fp(o), fp(d), u64p(ids), 256)
        if n == 0:
            continue
        ref = rng.random(4).astype(np.float32)
        for grey in (0, 1):
            a = Rf.R.ref_ray_loss(Rf.h, fp(ref), fp(o), fp(d), u64p(ids), n, grey)
            b = orc.orc_ray_loss(S.h, fp(ref), fp(o), fp(d), u64p(ids), n, grey)
            assert np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32)
        checked += 1
    assert checked > 50
    Rf.close()
    S.close()


@pytest.mark.skipif(not libs.have_ref("O0"), reason="needs the -O0 reference build")
def test_tv_value_matches_O0_reference(orc):
    """TVRegularisation (:196-248) is only meaningful in the as-shipped -O0 build (the SH loop at :241-245 is UB);
    there it returns exactly the density-only term the oracle restates."""
    gs = 8
    grid = libs.lcg_grid(gs, 20.0)
    Rf, S = RefScene(gs, grid, kind="O0", boxes=False), OracleScene(gs, grid)
    for z in range(gs - 1):
        for y in range(gs - 1):
            for x in range(gs - 1):
                a, b = Rf.R.ref_tv(Rf.h, x, y, z), orc.orc_tv(S.h, x, y, z)
                assert np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32), (x, y, z, a, b)
    Rf.close()
    S.close()


@pytest.mark.parametrize("which", [0, 1, 2])
def test_adam_bit_exact(orc, which):
    gs = 8
    grid = libs.lcg_grid(gs, 20.0)
    Rf, S = RefScene(gs, grid, boxes=False), OracleScene(gs, grid)
    rng = np.random.default_rng(which)
    for it in range(4):
        g = (rng.standard_normal((gs ** 3, 13)) * (rng.random((gs ** 3, 13)) < 0.4)).astype(np.float32)
        Rf.grad[:] = g
        S.grad[:] = g
        Rf.R.ref_adam_step(Rf.h, it, which)
        orc.orc_adam_step(S.h, it, which)
        assert np.array_equal(bits(Rf.grid), bits(S.grid))
        assert np.array_equal(bits(Rf.adam_v), bits(S.adam_v)) and np.array_equal(bits(Rf.adam_s), bits(S.adam_s))
    Rf.close()
    S.close()


def test_pack_color_and_libm(orc):
    R = libs.ref("O3")
    rng = np.random.default_rng(0)
    for c in np.concatenate([rng.random((200, 4)), [[0, 0, 0, 0], [1, 1, 1, 1], [0.999999, 0.5, 1.0 / 255, 1]]]).astype(np.float32):
        assert R.ref_pack_color(fp(c)) == orc.orc_pack_color(fp(c))
    # the glibc expf restatement the CUDA path carries (lrf_device.cuh expf_libm) == the libm the reference calls
    xs = -np.abs(rng.standard_normal(20000).astype(np.float32) * np.float32(3.0))
    xs = np.concatenate([xs, -rng.random(20000).astype(np.float32) * np.float32(1e-3), np.float32([-0.0, 0.0, -100.0, -87.5, -103.9, -104.5])])
    for x in xs:
        a, b = R.ref_expf(float(x)), orc.orc_expf_model(float(x), 1)
        assert np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32), (x, a, b)
