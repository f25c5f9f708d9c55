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


# ---- backward pin: directional finite differences of the COMPILED REFERENCE's own loss -----------------------------------------
# The reference's gradient is generated by Enzyme, which cannot be built here.  Its loss function can: RayGridLoss
# (example_tracer.cpp:258-268) and TVRegularisation (:196-248, -O0 object) are executed from oracle/_ref.  A central difference of
# the executed function along a direction v equals <grad, v> for ANY correct gradient — no restatement is involved on the
# reference side.  The fixture stores, per ray, (L(g + h v) - L(g - h v)) / 2h from the reference and the rays for which that
# quotient is a derivative at all: a ray whose shaded-sample count differs between g - h v, g and g + h v crossed the
# `throughput < 0.1` break (a jump of the loss, ref :131) and is left out.
BWD_CASES = {  # name -> (gs, density scale, view k, image size, greyscale, 1 / step)
    "opaque": (16, 20.0, 2, 24, False, 64),
    "thin": (16, 2.0, 5, 24, False, 64),
    "grey": (16, 20.0, 1, 20, True, 32),
}   # steps are powers of two (exact in fp32), chosen where rounding noise of the fp32 loss over 2h and curvature balance
BWD_DIRS = ("all", "density", "colour")


def fd_direction(gs, which, seed_k):
    """direction over the gs^3 x 13 parameters, reproducible from the repo's own LCG (no numpy RNG stream involved)"""
    u = libs.lcg_grid(gs, 1.0, seed=0x9E3779B97F4A7C15 ^ (0x1234567 * (seed_k + 1)))   # col 0: (u - .25), 1..3: u, 4: 1 + 2u, 5..: u - .5
    v = np.empty_like(u)
    v[:, 0] = (u[:, 0] + np.float32(0.25) - np.float32(0.5)) * np.float32(2.0)      # density: +-1
    v[:, 1:4] = (u[:, 1:4] - np.float32(0.5)) * np.float32(0.125)                 # RGB: +-1/16
    v[:, 4] = (u[:, 4] - np.float32(2.0)) * np.float32(0.0625)                     # sh0
    v[:, 5:] = u[:, 5:] * np.float32(0.125)                                        # sh1..8
    if which == "density":
        v[:, 1:] = 0
    elif which == "colour":
        v[:, 0] = 0
    return v


def backward_fd_case():
    from libs import u64p
    out = {"cases": np.array(list(BWD_CASES)), "dirs": np.array(BWD_DIRS)}
    for name, (gs, D, k, wh, grey, hdiv) in BWD_CASES.items():
        BWD_H = 1.0 / hdiv
        out[name + "_h"] = np.array([BWD_H])
        grid = libs.lcg_grid(gs, D)
        Rf = RefScene(gs, grid)
        vi, pi = Rf.set_cam(k, 7)
        n = wh * wh
        o, d = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
        hits = []
        ids = np.zeros(1024, np.uint64)
        for pix in range(n):
            Rf.R.ref_gen_ray(Rf.h, pix % wh, pix // wh, wh, wh, fp(o[pix]), fp(d[pix]))
            m = Rf.R.ref_ray_hits(Rf.h, fp(o[pix]), fp(d[pix]), u64p(ids), 1024)
            hits.append(ids[:m].copy())
        # target colours from the repo LCG as well
        big = libs.lcg_grid(8, 1.0, seed=0xABCDEF ^ (gs + k))[:, 1:4].reshape(-1)   # 1536 uniforms in [0, 1)
        ref = np.ones((n, 4), np.float32)
        ref[:, :3] = np.resize(big, n * 3).reshape(n, 3)
        if grey:   # kernel2D_RayMarchGrad mixes the target to luminance first (ref :696-704); equal channels make that the identity (to 1 ulp)
            ref[:, 1] = ref[:, 0]
            ref[:, 2] = ref[:, 0]
        _, ns0, _, _ = Rf.render_debug(wh, wh, grey, 0)
        out.update({name + "_gs": np.array([gs]), name + "_D": np.array([D]), name + "_wh": np.array([wh]), name + "_grey": np.array([int(grey)]),
                    name + "_view_inv": vi, name + "_proj_inv": pi, name + "_o": o, name + "_d": d, name + "_ref": ref,
                    name + "_n_hit": np.array([len(hh) for hh in hits], np.uint32)})

        def losses():
            L = np.zeros(n, np.float64)
            for pix in range(n):
                if hits[pix].size:
                    L[pix] = Rf.R.ref_ray_loss(Rf.h, fp(ref[pix]), fp(o[pix]), fp(d[pix]), u64p(hits[pix]), hits[pix].size, int(grey))
            return L

        L0 = losses()
        for di, which in enumerate(BWD_DIRS):
            if grey and which == "colour":
                continue
            v = fd_direction(gs, which, di)
            Rf.grid[:] = grid + np.float32(BWD_H) * v
            Lp = losses()
            _, nsp, _, _ = Rf.render_debug(wh, wh, grey, 0)
            Rf.grid[:] = grid - np.float32(BWD_H) * v
            Lm = losses()
            _, nsm, _, _ = Rf.render_debug(wh, wh, grey, 0)
            Rf.grid[:] = grid
            has = np.array([hh.size > 0 for hh in hits])
            # ... and a ray whose two one-sided quotients disagree crossed a kink of the loss inside [g - h v, g + h v] (a clamp of
            # ref :181-184, the relu :170, |C - ref| :265): there the central quotient is not the derivative at g either.  Both
            # filters use the reference's outputs only; they drop < 5 % of the rays.
            smooth = np.abs((Lp - L0) - (L0 - Lm)) <= 2e-6 + 0.01 * np.abs(Lp - Lm)
            kept = has & (nsp == ns0) & (nsm == ns0) & smooth
            out["%s_%s_kept" % (name, which)] = kept
            out["%s_%s_fd" % (name, which)] = np.where(kept, (Lp - Lm) / (2 * BWD_H), 0.0)
        Rf.close()
    if libs.have_ref("O0"):
        gs = 8
        grid = libs.lcg_grid(gs, 20.0)
        Rf = RefScene(gs, grid, kind="O0", boxes=False)
        v = fd_direction(gs, "density", 7)

        def tv_total():
            t = 0.0
            for z in range(gs - 1):
                for y in range(gs - 1):
                    for x in range(gs - 1):
                        t += float(Rf.R.ref_tv(Rf.h, x, y, z))
            return t

        hh = 1.0 / 256.0
        Rf.grid[:] = grid + np.float32(hh) * v
        tp = tv_total()
        Rf.grid[:] = grid - np.float32(hh) * v
        tm = tv_total()
        out.update({"tv_gs": np.array([gs]), "tv_h": np.array([hh]), "tv_fd": np.array([(tp - tm) / (2 * hh)])})
        Rf.close()
    return out


def bench_size_case():
    """the bench's own size: 256^3 dense grid, orbit views of the 100-view bench orbit, a 32 x 32 frame = every 25th pixel centre of
    the 800 x 800 frame ((i + .5) / 32 == (25 i + 12.5) / 800), opaque (D = 20) and thin (D = 2) density.  Octree build ~20 s."""
    gs, wh, keep = 256, 32, 12
    out = {"gs": gs, "wh": wh, "keep": keep, "views": np.array([0, 13, 37]), "n_views": 100, "Ds": np.array([20.0, 2.0])}
    Rf = RefScene(gs, None)
    for D in (20.0, 2.0):
        Rf.grid[:] = libs.lcg_grid(gs, D)
        for k in (0, 13, 37):
            vi, pi = Rf.set_cam(k, 100)
            nh, ns, rg, ids = Rf.render_debug(wh, wh, False, keep)
            img = np.zeros(wh * wh, np.uint32)
            Rf.R.ref_render(Rf.h, u32p(img), wh, wh, 0)
            tag = "D%d_v%d" % (int(D), k)
            out.update({"view_inv_" + tag: vi, "proj_inv_" + tag: pi, "n_hit_" + tag: nh.astype(np.uint16), "n_shaded_" + tag: ns.astype(np.uint16),
                        "rgbtc_" + tag: rg, "ids_" + tag: ids.astype(np.int32), "pixels_" + tag: img})
    Rf.close()
    return out


def main():
    assert libs.have_ref("O3"), "build oracle/_ref first: make -C oracle ref"
    os.makedirs(OUT, exist_ok=True)
    for name, args in FORWARD_CASES.items():
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **forward_case(*args))
    np.savez_compressed(os.path.join(OUT, "sparse_upsampled.npz"), **sparse_case())
    np.savez_compressed(os.path.join(OUT, "percell.npz"), **percell_case())
    np.savez_compressed(os.path.join(OUT, "io.npz"), **io_case())
    np.savez_compressed(os.path.join(OUT, "backward_fd.npz"), **backward_fd_case())
    np.savez_compressed(os.path.join(OUT, "bench256.npz"), **bench_size_case())
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
