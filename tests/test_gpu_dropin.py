"""Drop-in check: the reference's main.cpp, UNMODIFIED, linked against the replacement tracer class of
literf-radiancefields_b200/host/ (which forwards to liblrf_b200.so) runs its whole coarse-to-fine training
schedule on the GPU (ref main.cpp:74-134: 4 levels x 15 Adam steps x 7 views, 16^3 -> 128^3).

oracle/_ref/main_b200 is built in the dev container by `make -C oracle dropin` (needs /root/reference for main.cpp,
LiteMath, stb_image, nlohmann); the binary travels to the GPU box.  The test supplies what main.cpp reads from
disk: ../data/r_k.png (7 target views, written here as PNGs of renders of a synthetic blob scene made with our own
library) and the camera json main.cpp parses and never uses ("missing input")."""
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

import libs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "main_b200")


def write_png(path, rgb_top_down):
    h, w, _ = rgb_top_down.shape
    raw = b"".join(b"\x00" + rgb_top_down[y].tobytes() for y in range(h))

    def chunk(tag, data):
        c = struct.pack(">I", len(data)) + tag + data
        return c + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def read_bmp(path):
    """24-bit bottom-up BMP as written by LiteImage::SaveBMP -> [h, w, 3] RGB, row 0 = image bottom"""
    b = open(path, "rb").read()
    w, h = struct.unpack("<ii", b[18:26])
    px = np.frombuffer(b, np.uint8, w * h * 3, 54).reshape(h, w, 3)
    return px[:, :, ::-1]   # file order is B, G, R (SaveBMP puts bits 16-23 of the packed pixel first)


def unpack(pix, w):
    p = pix.reshape(-1, w)
    return np.stack([p & 255, (p >> 8) & 255, (p >> 16) & 255], axis=-1).astype(np.uint8)


@pytest.mark.gpu
def test_reference_main_runs_on_the_dropin_tracer(pkg, tmp_path):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/main_b200 not built (make -C oracle dropin, dev container only)")
    W = 256
    # ---- targets: an emissive blob rendered by our own library at main.cpp's 7 orbit views ----
    gs = 32
    g = libs.lcg_grid(gs, 40.0, blob=True)
    t = pkg.RayMarcherExample(0)
    t.InitGrid(gs)
    t.set_grid(g)
    t.SetWorldViewMProjatrix(pkg.perspective(45.0, 1.0, 0.1, 100.0))
    (tmp_path / "data" / "radiance_fields" / "lego").mkdir(parents=True)
    (tmp_path / "run").mkdir()
    (tmp_path / "data" / "radiance_fields" / "lego" / "train_camera_params.json").write_text("{}")
    targets = []
    for k in range(7):
        t.SetWorldViewMatrix(pkg.orbit_view(k, 7))
        pix = np.zeros(W * W, np.uint32)
        t.kernel2D_RayMarch(pix, W, W)
        rgb = unpack(pix, W)                      # row 0 = bottom
        # stbi_loadf linearises with gamma 2.2 (ref main.cpp:96, stb_image.h:1572): store the gamma-encoded value
        enc = np.round(255.0 * (rgb.astype(np.float64) / 255.0) ** (1 / 2.2)).astype(np.uint8)
        write_png(str(tmp_path / "data" / ("r_%d.png" % k)), enc[::-1])   # PNG rows are top-down; main.cpp flips on load (:72)
        targets.append((enc.astype(np.float64) / 255.0) ** 2.2)
    # what the untrained 16^3 grid of InitGrid renders to (the starting point of the schedule)
    t.InitGrid(16)
    init = []
    for k in range(7):
        t.SetWorldViewMatrix(pkg.orbit_view(k, 7))
        pix = np.zeros(W * W, np.uint32)
        t.kernel2D_RayMarch(pix, W, W)
        init.append(unpack(pix, W).astype(np.float64) / 255.0)
    t.close()
    # ---- the reference driver on the drop-in tracer ----
    r = subprocess.run([BIN], cwd=str(tmp_path / "run"), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Cleanup end" in r.stdout
    # main.cpp:136-137 prints "<boxes.size()> <(gs-1)^3>" with gs = 128 after three upsamples
    tail = [l for l in r.stdout.splitlines() if l.strip().endswith(str(127 ** 3))]
    assert tail, r.stdout[-500:]
    n_boxes = int(tail[-1].split()[0])
    assert 0 < n_boxes <= 127 ** 3
    # ---- the schedule must have trained: final 128^3 renders are far closer to the targets than the start ----
    mse_init = np.mean([(a - b) ** 2 for a, b in zip(init, targets)])
    finals = [read_bmp(str(tmp_path / "run" / ("out_cpu_128_%d.bmp" % k))).astype(np.float64) / 255.0 for k in range(7)]
    mse_final = np.mean([(a - b) ** 2 for a, b in zip(finals, targets)])
    psnr = 10 * np.log10(1.0 / max(mse_final, 1e-12))
    print("drop-in main.cpp: boxes %d / %d, MSE start %.4f -> final %.5f (PSNR %.1f dB)" % (n_boxes, 127 ** 3, mse_init, mse_final, psnr))
    assert mse_final < 0.25 * mse_init
    for lvl in (16, 32, 64, 128):
        assert os.path.exists(str(tmp_path / "run" / ("out_cpu_%d_0.bmp" % lvl)))
