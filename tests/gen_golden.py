#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/libref_O3.so and, for the TV
value, libref_O0.so — compiled from /root/reference by oracle/Makefile) on seeded synthetic inputs.

    python tests/gen_golden.py          # needs /root/reference (dev container only); rewrites tests/golden/

The inputs are not stored: grids come from the seeded LCG of SURVEY 8(d) (libs.lcg_grid) and cameras from the
reference's own caller-side formulas (main.cpp:78,99-102); the fixtures hold the cameras' inverted matrices
as the reference stored them plus every per-ray output.  tests/test_golden.py replays them against the oracle
(CPU) and against the CUDA path (GPU) — on the GPU box /root/reference does not exist, the fixtures do.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import libs
from libs import fp, u32p
from test_oracle_vs_ref import RefScene

OUT = os.path.join(HERE, "golden")

# name -> (gs, density scale, blob, image size, views (k of 7), greyscale, ids kept per ray)
FORWARD_CASES = {
    "fwd_gs16_opaque": (16, 20.0, False, 32, (0, 3, 5), False, 40),
    "fwd_gs16_thin": (16, 2.0, False, 32, (1, 6), False, 40),
    "fwd_gs16_grey": (16, 20.0, False, 24, (2,), True, 40),
    "fwd_gs32_opaque": (32, 20.0, False, 32, (0, 4), False, 80),
    "fwd_gs64_blob": (64, 20.0, True, 24, (1,), False, 150),
}


def forward_case(gs, D, blob, wh, views, grey, keep):
    grid = libs.lcg_grid(gs, D, blob=blob)
    Rf = RefScene(gs, grid)
    out = {"gs": gs, "D": D, "blob": int(blob), "wh": wh, "views": np.array(views), "grey": int(grey), "keep": keep}
    for k in views:
        vi, pi = Rf.set_cam(k, 7)
        nh, ns, rg, ids = Rf.render_debug(wh, wh, grey, keep)
        img = np.zeros(wh * wh, np.uint32)
        Rf.R.ref_render(Rf.h, u32p(img), wh, wh, int(grey))
        out.update({"view_inv_%d" % k: vi, "proj_inv_%d" % k: pi, "n_hit_%d" % k: nh.astype(np.uint16), "n_shaded_%d" % k: ns.astype(np.uint16),
                    "rgbtc_%d" % k: rg, "ids_%d" % k: ids.astype(np.int32), "pixels_%d" % k: img})
    Rf.close()
    return out


def sparse_case():
    """8^3 blob -> 2 x (UpsampleGrid, UpsampleOctree) -> 32^3 sparse set in the reference's own id order, rendered"""
    gs = 8
    grid = libs.lcg_grid(gs, 20.0, blob=True)
    grid[:, 0] -= 0.4
    Rf = RefScene(gs, grid)
    out = {"gs0": gs}
    for level in range(2):
        Rf.R.ref_upsample_grid(Rf.h)
        Rf.R.ref_upsample_octree(Rf.h)
        out["boxes_%d" % level] = Rf.boxes_xyz().astype(np.int16)
        out["grid_sum_%d" % level] = np.array([Rf.grid.astype(np.float64).sum()])
        out["grid_sample_%d" % level] = Rf.grid[:: max(1, Rf.grid.shape[0] // 512)].copy()
        vi, pi = Rf.set_cam(1 + level, 7)
        nh, ns, rg, ids = Rf.render_debug(40, 40, False, 40)
        out.update({"view_inv_%d" % level: vi, "proj_inv_%d" % level: pi, "n_hit_%d" % level: nh.astype(np.uint16),
                    "n_shaded_%d" % level: ns.astype(np.uint16), "rgbtc_%d" % level: rg, "ids_%d" % level: ids.astype(np.int32)})
    Rf.close()
    return out


def percell_case():
    """Adam (3 variants, 3 iterations each) and the TV value on an 8^3 grid"""
    gs = 8
    grid = libs.lcg_grid(gs, 20.0)
    out = {"gs": gs}
    for which in (0, 1, 2):
        Rf = RefScene(gs, grid, boxes=False)
        rng = np.random.default_rng(100 + which)
        for it in range(3):
            g = (rng.standard_normal((gs ** 3, 13)) * (rng.random((gs ** 3, 13)) < 0.4)).astype(np.float32)
            Rf.grad[:] = g
            Rf.R.ref_adam_step(Rf.h, it, which)
        out["adam_p_%d" % which] = Rf.grid.copy()
        out["adam_v_%d" % which] = Rf.adam_v.copy()
        out["adam_s_%d" % which] = Rf.adam_s.copy()
        Rf.close()
    if libs.have_ref("O0"):
        Rf = RefScene(gs, grid, kind="O0", boxes=False)
        tv = np.zeros((gs - 1, gs - 1, gs - 1), np.float32)
        for z in range(gs - 1):
            for y in range(gs - 1):
                for x in range(gs - 1):
                    tv[z, y, x] = Rf.R.ref_tv(Rf.h, x, y, z)
        out["tv_O0"] = tv
        Rf.close()
    return out


def io_case():
    """a seeded 37x23 RGBA8 image through the reference's own stbi_loadf (flip on, as main.cpp:72 sets it) and a packed
    frame through its LiteImage::SaveBMP (example_tracer.cpp:614)"""
    import ctypes as C
    import tempfile
    R = libs.ref("O3")
    rng = np.random.default_rng(7)
    w, h = 37, 23
    rgba8 = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    rgba8[0, :8] = np.arange(8, dtype=np.uint8)[:, None] * 36      # make sure 0, 255 and small values occur
    rgba8[1, :4] = 255
    out = {"rgba8": rgba8}
    with tempfile.TemporaryDirectory() as td:
        png = os.path.join(td, "t.png")
        libs.write_png_rgba8(png, rgba8)
        for flip in (0, 1):
            f = np.zeros(w * h * 4, np.float32)
            ww, hh = C.c_int(0), C.c_int(0)
            assert R.ref_stbi_loadf(png.encode(), flip, fp(f), f.size, C.byref(ww), C.byref(hh)) == 0 and (ww.value, hh.value) == (w, h)
            out["loadf_flip%d" % flip] = f.reshape(h * w, 4)
        pix = rng.integers(0, 2 ** 32, 29 * 17, dtype=np.uint64).astype(np.uint32)
        bmp = os.path.join(td, "t.bmp")
        assert R.ref_save_bmp(bmp.encode(), u32p(pix), 29, 17) == 0
        out["bmp_pixels"] = pix
        out["bmp_file"] = np.frombuffer(open(bmp, "rb").read(), dtype=np.uint8)
    return out


def main():
    assert libs.have_ref("O3"), "build oracle/_ref first: make -C oracle ref"
    os.makedirs(OUT, exist_ok=True)
    for name, args in FORWARD_CASES.items():
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **forward_case(*args))
    np.savez_compressed(os.path.join(OUT, "sparse_upsampled.npz"), **sparse_case())
    np.savez_compressed(os.path.join(OUT, "percell.npz"), **percell_case())
    np.savez_compressed(os.path.join(OUT, "io.npz"), **io_case())
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
