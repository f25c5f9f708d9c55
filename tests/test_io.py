"""Callers either side of the path: model.dat, target ingest (stbi_loadf semantics), BMP output.
CPU part: the oracle's restatements against the compiled reference (stb_image.h / Image2d.cpp where they lie) and
against the committed fixture tests/golden/io.npz; the C-ABI's host-only BMP writer.  GPU part: the device ingest,
batch gather, resident-target backward, model round trip and the rendered BMP, through the C-ABI."""
import ctypes as C
import os

import numpy as np
import pytest

import libs
from libs import fp, u32p

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "io.npz")


def oracle_loadf(orc, rgba8, flip):
    h, w, _ = rgba8.shape
    out = np.zeros((h * w, 4), np.float32)
    a = np.ascontiguousarray(rgba8)
    orc.orc_ldr_to_hdr_rgba8(a.ctypes.data, w, h, int(flip), fp(out))
    return out


def oracle_bmp(orc, pix, w, h):
    buf = np.zeros(54 + 3 * w * h, np.uint8)
    n = orc.orc_bmp_bytes(u32p(pix), w, h, buf.ctypes.data)
    assert n == buf.size
    return buf


def test_oracle_io_matches_golden(orc):
    g = np.load(GOLD)
    for flip in (0, 1):
        assert np.array_equal(oracle_loadf(orc, g["rgba8"], flip).view(np.uint32), g["loadf_flip%d" % flip].view(np.uint32))
    assert np.array_equal(oracle_bmp(orc, g["bmp_pixels"], 29, 17), g["bmp_file"])


@pytest.mark.skipif(not libs.have_ref("O3"), reason="compiled reference (oracle/_ref) not present")
def test_oracle_io_matches_reference(orc, tmp_path):
    R = libs.ref("O3")
    rng = np.random.default_rng(11)
    for (w, h) in ((1, 1), (5, 3), (64, 48)):
        rgba8 = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        png = str(tmp_path / "t.png")
        libs.write_png_rgba8(png, rgba8)
        for flip in (0, 1):
            f = np.zeros(w * h * 4, np.float32)
            ww, hh = C.c_int(0), C.c_int(0)
            assert R.ref_stbi_loadf(png.encode(), flip, fp(f), f.size, C.byref(ww), C.byref(hh)) == 0
            assert (ww.value, hh.value) == (w, h)
            assert np.array_equal(f.view(np.uint32), oracle_loadf(orc, rgba8, flip).ravel().view(np.uint32))
        pix = rng.integers(0, 2 ** 32, w * h, dtype=np.uint64).astype(np.uint32)
        bmp = str(tmp_path / "t.bmp")
        assert R.ref_save_bmp(bmp.encode(), u32p(pix), w, h) == 0
        assert np.array_equal(np.frombuffer(open(bmp, "rb").read(), np.uint8), oracle_bmp(orc, pix, w, h))


def test_capi_save_bmp_bytes(pkg, tmp_path):
    """lrf_save_bmp is host code (no device call): byte-identical to the reference's file (fixture) and empty-image edge"""
    g = np.load(GOLD)
    path = str(tmp_path / "o.bmp")
    pkg.save_bmp(path, g["bmp_pixels"], 29, 17)
    assert np.array_equal(np.frombuffer(open(path, "rb").read(), np.uint8), g["bmp_file"])
    pkg.save_bmp(path, np.zeros(1, np.uint32), 0, 0)
    assert os.path.getsize(path) == 54


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture()
def tracer(pkg):
    t = pkg.RayMarcherExample(0)
    yield t
    t.close()


@pytest.mark.gpu
def test_gpu_target_ingest_bit_exact(pkg, orc, tracer):
    g = np.load(GOLD)
    rgba8 = g["rgba8"]
    h, w, _ = rgba8.shape
    tracer.targets_resize(3, w, h)
    tracer.target_ingest_rgba8(0, rgba8, flip_vertical=True)
    tracer.target_ingest_rgba8(2, rgba8, flip_vertical=False)
    assert np.array_equal(tracer.target_download(0, w, h).view(np.uint32), g["loadf_flip1"].view(np.uint32))
    assert np.array_equal(tracer.target_download(2, w, h).view(np.uint32), g["loadf_flip0"].view(np.uint32))
    # every 8-bit value in every channel, against the oracle
    allv = np.zeros((4, 256, 4), np.uint8)
    allv[:] = np.arange(256, dtype=np.uint8)[None, :, None]
    tracer.targets_resize(1, 256, 4)
    tracer.target_ingest_rgba8(0, allv, flip_vertical=True)
    assert np.array_equal(tracer.target_download(0, 256, 4).view(np.uint32), oracle_loadf(orc, allv, 1).view(np.uint32))
    with pytest.raises(pkg.LrfError):
        tracer.target_ingest_rgba8(5, allv)


@pytest.mark.gpu
def test_gpu_target_gather_and_resident_backward(pkg, orc, tracer):
    from scenes import cameras
    gs, w = 16, 24
    grid = libs.lcg_grid(gs, 20.0)
    tracer.InitGrid(gs)
    tracer.set_grid(grid)
    vi, pi = cameras(None, 2, 7)
    tracer.set_view_inv(vi)
    tracer.set_proj_inv(pi)
    rng = np.random.default_rng(3)
    rgba8 = rng.integers(0, 256, (2, w, w, 4), dtype=np.uint8)
    tracer.targets_resize(2, w, w)
    for v in range(2):
        tracer.target_ingest_rgba8(v, rgba8[v], flip_vertical=True)
    ref1 = oracle_loadf(orc, rgba8[1], 1)
    # gather: picks {view, pixel}
    n = 1000
    picks = np.stack([rng.integers(0, 2, n), rng.integers(0, w * w, n)], axis=1).astype(np.uint32)
    pb, ob = pkg.DeviceBuffer(tracer, picks.nbytes), pkg.DeviceBuffer(tracer, n * 16)
    pb.upload(picks)
    tracer.gather_targets_dev_async(n, pb.ptr, ob.ptr)
    got = ob.download(np.zeros((n, 4), np.float32))
    both = np.stack([oracle_loadf(orc, rgba8[0], 1), ref1])
    assert np.array_equal(got.view(np.uint32), both[picks[:, 0], picks[:, 1]].view(np.uint32))
    # backward against the resident target == backward against the same pixels handed over from the host
    tracer.zeroGrad()
    tracer.render_grad_target_async(1)
    tracer.sync()
    g_res = tracer.get_grad().astype(np.float64)
    tracer.zeroGrad()
    tracer.kernel2D_RayMarchGrad(w, w, ref1)
    g_host = tracer.get_grad().astype(np.float64)
    assert np.linalg.norm(g_res) > 0
    assert np.linalg.norm(g_res - g_host) <= 1e-5 * np.linalg.norm(g_host)   # same kernel, atomics in another order


@pytest.mark.gpu
def test_gpu_model_dat_round_trip(pkg, tracer, tmp_path):
    gs = 16
    grid = libs.lcg_grid(gs, 20.0)
    tracer.InitGrid(gs)
    tracer.set_grid(grid)
    path = str(tmp_path / "model.dat")
    tracer.save_model(path)
    # the file is the reference's raw dump: gs^3 cells x 13 floats, AoS (main.cpp:139-141)
    assert os.path.getsize(path) == gs ** 3 * 52
    assert np.array_equal(np.fromfile(path, np.float32).view(np.uint32), grid.ravel().view(np.uint32))
    tracer.InitGrid(gs)
    tracer.load_model(path)
    assert np.array_equal(tracer.get_grid().view(np.uint32), grid.view(np.uint32))
    tracer.InitGrid(8)
    with pytest.raises(pkg.LrfError):
        tracer.load_model(path)          # size mismatch is an error, not a partial read
    with pytest.raises(pkg.LrfError):
        tracer.load_model(str(tmp_path / "absent.dat"))


@pytest.mark.gpu
def test_gpu_render_bmp_matches_reference_writer(pkg, orc, tracer, tmp_path):
    from scenes import cameras
    gs, w, h = 16, 40, 24
    tracer.InitGrid(gs)
    tracer.set_grid(libs.lcg_grid(gs, 20.0))
    vi, pi = cameras(None, 4, 7)
    tracer.set_view_inv(vi)
    tracer.set_proj_inv(pi)
    frame = np.zeros(w * h, np.uint32)
    tracer.kernel2D_RayMarch(frame, w, h)
    path = str(tmp_path / "out.bmp")
    tracer.render_bmp(path, w, h)
    assert np.array_equal(np.frombuffer(open(path, "rb").read(), np.uint8), oracle_bmp(orc, frame, w, h))
