"""The backward, pinned against numbers the UNMODIFIED reference produced (tests/golden/backward_fd.npz, tests/gen_golden.py).

The reference's gradient comes out of the Enzyme clang plugin (example_tracer.cpp:270-279, :250-256), which cannot be built here —
but the functions Enzyme differentiates can: RayGridLoss (ref :258-268) and TVRegularisation (ref :196-248, -O0 object) are executed
from oracle/_ref.  A central difference of the executed function along a direction v is <grad, v> of any correct gradient, whatever
produced it.  The fixture holds that quotient per ray (and for the TV sum), for three directions (all 13 channels / density only /
colour + SH only), over the rays for which it is a derivative at all (no change of the shaded-sample count, no kink inside the
step: both decided from the reference's own outputs).  Checked here:
  * not gpu: the oracle's hand-derived adjoint (what every CUDA gradient test compares with)   -> moves a11 / a12 from
    "self-certified" to pinned against executed reference code;
  * gpu: the CUDA packed-ray backward and the CUDA TV gradient through the C-ABI, against the same reference numbers.
Tolerance (BASELINE.json north_star: gradients within 1e-3 relative): the vector of per-ray directional derivatives agrees with
the reference's to 1e-3 in the 2-norm; measured 2e-4 .. 9e-4, the floor being the fp32 rounding of the reference's loss over 2h."""
import os

import numpy as np
import pytest

import libs
from libs import fp, u64p
from scenes import OracleScene

TOL = 1e-3
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "backward_fd.npz")


def fd_direction(gs, which, seed_k):
    """same construction as tests/gen_golden.py::fd_direction (kept here so that the GPU box needs no generator import)"""
    u = libs.lcg_grid(gs, 1.0, seed=0x9E3779B97F4A7C15 ^ (0x1234567 * (seed_k + 1)))
    v = np.empty_like(u)
    v[:, 0] = (u[:, 0] + np.float32(0.25) - np.float32(0.5)) * np.float32(2.0)
    v[:, 1:4] = (u[:, 1:4] - np.float32(0.5)) * np.float32(0.125)
    v[:, 4] = (u[:, 4] - np.float32(2.0)) * np.float32(0.0625)
    v[:, 5:] = u[:, 5:] * np.float32(0.125)
    if which == "density":
        v[:, 1:] = 0
    elif which == "colour":
        v[:, 0] = 0
    return v


def _cases():
    F = np.load(GOLD)
    for name in F["cases"]:
        for di, which in enumerate(F["dirs"]):
            if "%s_%s_fd" % (name, which) in F:
                yield str(name), di, str(which)


@pytest.mark.parametrize("name,di,which", list(_cases()))
def test_oracle_adjoint_against_reference_fd(orc, name, di, which):
    F = np.load(GOLD)
    gs, D, wh, grey = int(F[name + "_gs"][0]), float(F[name + "_D"][0]), int(F[name + "_wh"][0]), int(F[name + "_grey"][0])
    S = OracleScene(gs, libs.lcg_grid(gs, D))
    S.set_cam(F[name + "_view_inv"], F[name + "_proj_inv"])
    o, d, ref = F[name + "_o"], F[name + "_d"], F[name + "_ref"]
    v = fd_direction(gs, which, di).astype(np.float64)
    kept, fd = F["%s_%s_kept" % (name, which)], F["%s_%s_fd" % (name, which)]
    assert kept.sum() >= 0.9 * (F[name + "_n_hit"] > 0).sum()       # the filters drop only a few rays
    an = np.zeros(wh * wh)
    ids = np.zeros(1024, np.uint64)
    for pix in np.nonzero(kept)[0]:
        m = orc.orc_ray_hits(S.h, fp(o[pix]), fp(d[pix]), u64p(ids), 1024)
        assert m == F[name + "_n_hit"][pix]
        S.grad[:] = 0
        orc.orc_ray_grad(S.h, fp(ref[pix]), fp(o[pix]), fp(d[pix]), u64p(ids), m, grey)
        an[pix] = (S.grad.astype(np.float64) * v).sum()
    err = np.linalg.norm((an - fd)[kept]) / np.linalg.norm(fd[kept])
    assert err <= TOL, "oracle adjoint vs reference finite differences: %.2e" % err
    S.close()


def test_oracle_tv_gradient_against_reference_fd(orc):
    F = np.load(GOLD)
    if "tv_fd" not in F:
        pytest.skip("fixture generated without the -O0 reference build")
    gs = int(F["tv_gs"][0])
    S = OracleScene(gs, libs.lcg_grid(gs, 20.0))
    S.grad[:] = 0
    orc.orc_tv_grad(S.h)
    an = (S.grad.astype(np.float64) * fd_direction(gs, "density", 7)).sum()
    assert abs(an - F["tv_fd"][0]) <= TOL * abs(F["tv_fd"][0])
    S.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,di,which", list(_cases()))
def test_cuda_gradient_against_reference_fd(pkg, name, di, which):
    F = np.load(GOLD)
    gs, D, wh, grey = int(F[name + "_gs"][0]), float(F[name + "_D"][0]), int(F[name + "_wh"][0]), int(F[name + "_grey"][0])
    t = pkg.RayMarcherExample()
    t.InitGrid(gs)
    t.SetBoundingBox([0, 0, 0], [1, 1, 1])
    t.set_grid(libs.lcg_grid(gs, D))
    o, d, ref = F[name + "_o"], F[name + "_d"], F[name + "_ref"]
    v = fd_direction(gs, which, di).astype(np.float64)
    kept, fd = F["%s_%s_kept" % (name, which)], F["%s_%s_fd" % (name, which)]
    an = np.zeros(wh * wh)
    for pix in np.nonzero(kept)[0]:
        t.zeroGrad()
        t.rays_grad(o[pix:pix + 1], d[pix:pix + 1], ref[pix:pix + 1], bool(grey))
        an[pix] = (t.get_grad().astype(np.float64) * v).sum()
    err = np.linalg.norm((an - fd)[kept]) / np.linalg.norm(fd[kept])
    assert err <= TOL, "CUDA gradient vs reference finite differences: %.2e" % err
    # the whole batch in one call (the path training takes): sum over the kept rays
    t.zeroGrad()
    t.rays_grad(o[kept], d[kept], ref[kept], bool(grey))
    tot = (t.get_grad().astype(np.float64) * v).sum()
    assert abs(tot - an[kept].sum()) <= 1e-4 * np.abs(an[kept]).sum()
    t.close()


@pytest.mark.gpu
def test_cuda_tv_gradient_against_reference_fd(pkg):
    F = np.load(GOLD)
    if "tv_fd" not in F:
        pytest.skip("fixture generated without the -O0 reference build")
    gs = int(F["tv_gs"][0])
    t = pkg.RayMarcherExample()
    t.InitGrid(gs)
    t.set_grid(libs.lcg_grid(gs, 20.0))
    t.zeroGrad()
    t.kernel2D_TVGrad()
    an = (t.get_grad().astype(np.float64) * fd_direction(gs, "density", 7)).sum()
    assert abs(an - F["tv_fd"][0]) <= TOL * abs(F["tv_fd"][0])
    t.close()
