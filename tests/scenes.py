"""Shared scene builders for the parity tests: the same seeded inputs go to the oracle and to the CUDA path."""
import ctypes as C

import numpy as np

import libs
from libs import fp, i32p, i64p, u32p, u64p, np_view


def cameras(pkg_or_none, k, n_views, fov=45.0, aspect=1.0):
    """(view_inv, proj_inv) of main.cpp:78,99-104 through the ORACLE's LiteMath restatement"""
    O = libs.oracle()
    proj, view = np.zeros(16, np.float32), np.zeros(16, np.float32)
    O.orc_perspective(fov, aspect, 0.1, 100.0, fp(proj))
    O.orc_orbit_view(k, n_views, fp(view))
    vi, pi = np.zeros(16, np.float32), np.zeros(16, np.float32)
    O.orc_inverse4x4(fp(view), fp(vi))
    O.orc_inverse4x4(fp(proj), fp(pi))
    return vi, pi


class OracleScene:
    def __init__(self, gs, grid=None):
        self.O = libs.oracle()
        self.h = self.O.orc_create(gs)
        self.gs = gs
        if grid is not None:
            self.grid[:] = np.asarray(grid, np.float32).reshape(-1, 13)

    @property
    def grid(self):
        return np_view(self.O.orc_grid_ptr(self.h), self.O.orc_grid_size(self.h) ** 3 * 13).reshape(-1, 13)

    @property
    def grad(self):
        return np_view(self.O.orc_grad_ptr(self.h), self.O.orc_grid_size(self.h) ** 3 * 13).reshape(-1, 13)

    @property
    def adam_v(self):
        return np_view(self.O.orc_adam_v_ptr(self.h), self.O.orc_grid_size(self.h) ** 3 * 13).reshape(-1, 13)

    @property
    def adam_s(self):
        return np_view(self.O.orc_adam_s_ptr(self.h), self.O.orc_grid_size(self.h) ** 3 * 13).reshape(-1, 13)

    def set_cam(self, vi, pi):
        self.O.orc_set_view_inv(self.h, fp(vi))
        self.O.orc_set_proj_inv(self.h, fp(pi))

    def boxes(self):
        out = np.zeros((self.O.orc_box_count(self.h), 3), np.int32)
        self.O.orc_get_boxes_xyz(self.h, i32p(out))
        return out

    def set_boxes(self, xyz):
        xyz = np.ascontiguousarray(xyz, np.int32).reshape(-1, 3)
        self.O.orc_set_boxes_xyz(self.h, i32p(xyz), xyz.shape[0])

    def render(self, w, h, grey=False):
        out = np.zeros(w * h, np.uint32)
        self.O.orc_render(self.h, u32p(out), w, h, int(grey))
        return out

    def render_debug(self, w, h, grey=False, keep=0):
        n = w * h
        nh, ns = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        rg = np.zeros((n, 5), np.float32)
        ids = np.zeros((n, max(keep, 1)), np.int64)
        self.O.orc_render_debug(self.h, w, h, int(grey), u32p(nh), u32p(ns), fp(rg), i64p(ids), keep)
        vox = None
        if keep:
            b = self.boxes()
            vox = np.where(ids[..., None] >= 0, b[np.maximum(ids, 0)], -1).astype(np.int32)
        return nh, ns, rg, vox

    def trace_rays(self, o, d, grey=False, keep=0):
        o = np.ascontiguousarray(o, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
        n = o.shape[0]
        nh, ns = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        rg = np.zeros((n, 5), np.float32)
        vox = np.zeros((n, max(keep, 1), 3), np.int32)
        self.O.orc_trace_rays(self.h, n, fp(o), fp(d), int(grey), u32p(nh), u32p(ns), fp(rg), i32p(vox), keep)
        return nh, ns, rg, (vox if keep else None)

    def gen_rays(self, w, h):
        o, d = np.zeros((w * h, 3), np.float32), np.zeros((w * h, 3), np.float32)
        oo, dd = np.zeros(3, np.float32), np.zeros(3, np.float32)
        for y in range(h):
            for x in range(w):
                self.O.orc_gen_ray(self.h, x, y, w, h, fp(oo), fp(dd))
                o[y * w + x], d[y * w + x] = oo, dd
        return o, d

    def close(self):
        if self.h:
            self.O.orc_destroy(self.h)
            self.h = None


def sphere_boxes(gs, radius=0.288):
    """config-4 style sparse set: voxels whose centre lies inside a sphere, lattice-linear id order"""
    n = gs - 1
    idx = np.arange(n)
    c = (idx + 0.5) / gs - 0.5
    z, y, x = np.meshgrid(c, c, c, indexing="ij")
    m = (x * x + y * y + z * z) < radius * radius
    zz, yy, xx = np.nonzero(m)
    return np.stack([xx, yy, zz], axis=1).astype(np.int32)
