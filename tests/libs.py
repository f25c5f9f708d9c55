"""ctypes loaders for the test-side libraries: the CPU oracle (oracle/_build) and, where it was
built (this container only), the compiled reference (oracle/_ref).  TEST INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "liblrf_oracle.so")
REF_SO = {"O3": os.path.join(ROOT, "oracle", "_ref", "libref_O3.so"), "O0": os.path.join(ROOT, "oracle", "_ref", "libref_O0.so")}

c_fp = C.POINTER(C.c_float)
c_dp = C.POINTER(C.c_double)
c_u32p = C.POINTER(C.c_uint32)
c_u64p = C.POINTER(C.c_uint64)
c_i64p = C.POINTER(C.c_int64)
c_i32p = C.POINTER(C.c_int32)


def fp(a):
    return a.ctypes.data_as(c_fp)


def dp(a):
    return a.ctypes.data_as(c_dp)


def u32p(a):
    return a.ctypes.data_as(c_u32p)


def u64p(a):
    return a.ctypes.data_as(c_u64p)


def i64p(a):
    return a.ctypes.data_as(c_i64p)


def i32p(a):
    return a.ctypes.data_as(c_i32p)


def build_oracle():
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(ROOT, "oracle", "lrf_oracle.cpp")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], stdout=subprocess.DEVNULL)


_oracle = None


def oracle():
    global _oracle
    if _oracle is not None:
        return _oracle
    build_oracle()
    L = C.CDLL(ORACLE_SO)
    vp = C.c_void_p
    sigs = {
        "orc_perspective": (None, [C.c_float] * 4 + [c_fp]),
        "orc_orbit_view": (None, [C.c_int, C.c_int, c_fp]),
        "orc_inverse4x4": (None, [c_fp, c_fp]),
        "orc_create": (vp, [C.c_int]),
        "orc_destroy": (None, [vp]),
        "orc_grid_size": (C.c_int, [vp]),
        "orc_grid_ptr": (c_fp, [vp]),
        "orc_grad_ptr": (c_fp, [vp]),
        "orc_adam_v_ptr": (c_fp, [vp]),
        "orc_adam_s_ptr": (c_fp, [vp]),
        "orc_box_count": (C.c_size_t, [vp]),
        "orc_get_boxes_xyz": (None, [vp, c_i32p]),
        "orc_set_boxes_xyz": (None, [vp, c_i32p, C.c_size_t]),
        "orc_set_bbox": (None, [vp, c_fp, c_fp]),
        "orc_set_view_inv": (None, [vp, c_fp]),
        "orc_set_proj_inv": (None, [vp, c_fp]),
        "orc_gen_ray": (None, [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, c_fp, c_fp]),
        "orc_ray_hits": (C.c_size_t, [vp, c_fp, c_fp, c_u64p, C.c_size_t]),
        "orc_ray_hits_generic": (C.c_size_t, [C.c_int, c_dp, C.c_size_t, c_dp, c_dp, c_u64p, C.c_size_t]),
        "orc_shade": (C.c_size_t, [vp, c_fp, c_fp, c_u64p, C.c_size_t, C.c_int, c_fp, c_fp, c_fp]),
        "orc_render": (None, [vp, c_u32p, C.c_uint32, C.c_uint32, C.c_int]),
        "orc_render_debug": (None, [vp, C.c_uint32, C.c_uint32, C.c_int, c_u32p, c_u32p, c_fp, c_i64p, C.c_uint32]),
        "orc_trace_rays": (None, [vp, C.c_size_t, c_fp, c_fp, C.c_int, c_u32p, c_u32p, c_fp, c_i32p, C.c_uint32]),
        "orc_count_touched_nodes": (C.c_uint64, [vp, C.c_uint32, C.c_uint32, C.c_int]),
        "orc_ray_loss": (C.c_float, [vp, c_fp, c_fp, c_fp, c_u64p, C.c_size_t, C.c_int]),
        "orc_ray_loss_f64": (C.c_double, [vp, c_dp, c_fp, c_fp, c_fp, c_u64p, C.c_size_t, C.c_int]),
        "orc_ray_grad": (None, [vp, c_fp, c_fp, c_fp, c_u64p, C.c_size_t, C.c_int]),
        "orc_ray_grad_f64": (None, [vp, c_dp, c_dp, c_fp, c_fp, c_fp, c_u64p, C.c_size_t, C.c_int]),
        "orc_render_grad": (None, [vp, C.c_uint32, C.c_uint32, c_fp, C.c_int]),
        "orc_rays_grad": (None, [vp, C.c_size_t, c_fp, c_fp, c_fp, C.c_int]),
        "orc_tv": (C.c_float, [vp, C.c_size_t, C.c_size_t, C.c_size_t]),
        "orc_tv_total_f64": (C.c_double, [vp, c_dp]),
        "orc_tv_grad": (None, [vp]),
        "orc_zero_grad": (None, [vp]),
        "orc_adam_init": (None, [vp]),
        "orc_adam_step": (None, [vp, C.c_int, C.c_int]),
        "orc_upsample_grid": (None, [vp]),
        "orc_upsample_boxes": (None, [vp]),
        "orc_rebuild_boxes": (None, [vp]),
        "orc_pack_color": (C.c_uint32, [c_fp]),
        "orc_expf_model": (C.c_float, [C.c_float, C.c_int]),
        "orc_lcg_fill_grid": (None, [vp, C.c_float, C.c_int]),
        "orc_omp_threads": (C.c_int, []),
        "orc_set_omp_threads": (None, [C.c_int]),
        "orc_ldr_to_hdr_rgba8": (None, [vp, C.c_uint32, C.c_uint32, C.c_int, c_fp]),
        "orc_bmp_bytes": (C.c_size_t, [c_u32p, C.c_int, C.c_int, vp]),
    }
    for name, (res, args) in sigs.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _oracle = L
    return L


_ref = {}


def have_ref(kind="O3"):
    return os.path.exists(REF_SO[kind])


def ref(kind="O3"):
    if kind in _ref:
        return _ref[kind]
    L = C.CDLL(REF_SO[kind])
    vp = C.c_void_p
    sigs = {
        "ref_create": (vp, []),
        "ref_destroy": (None, [vp]),
        "ref_init_grid": (None, [vp, C.c_int]),
        "ref_init_grid_nobox": (None, [vp, C.c_int]),
        "ref_set_bbox": (None, [vp, c_fp, c_fp]),
        "ref_grid_size": (C.c_size_t, [vp]),
        "ref_grid_ptr": (c_fp, [vp]),
        "ref_grad_ptr": (c_fp, [vp]),
        "ref_adam_v_ptr": (c_fp, [vp]),
        "ref_adam_s_ptr": (c_fp, [vp]),
        "ref_get_adam_hparams": (None, [vp, c_fp]),
        "ref_box_count": (C.c_size_t, [vp]),
        "ref_get_boxes": (None, [vp, c_dp]),
        "ref_set_boxes": (None, [vp, c_dp, C.c_size_t]),
        "ref_make_perspective": (None, [C.c_float] * 4 + [c_fp]),
        "ref_make_orbit_view": (None, [C.c_int, C.c_int, c_fp]),
        "ref_inverse4x4": (None, [c_fp, c_fp]),
        "ref_set_view_matrix": (None, [vp, c_fp]),
        "ref_set_proj_matrix": (None, [vp, c_fp]),
        "ref_get_inv_matrices": (None, [vp, c_fp, c_fp]),
        "ref_render": (None, [vp, c_u32p, C.c_uint32, C.c_uint32, C.c_int]),
        "ref_gen_ray": (None, [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, c_fp, c_fp]),
        "ref_ray_hits": (C.c_size_t, [vp, c_fp, c_fp, c_u64p, C.c_size_t]),
        "ref_shade": (None, [vp, c_fp, c_fp, c_u64p, C.c_size_t, C.c_int, c_fp, c_fp, c_fp]),
        "ref_render_debug": (None, [vp, C.c_uint32, C.c_uint32, C.c_int, c_u32p, c_u32p, c_fp, c_i64p, C.c_uint32]),
        "ref_ray_loss": (C.c_float, [vp, c_fp, c_fp, c_fp, c_u64p, C.c_size_t, C.c_int]),
        "ref_tv": (C.c_float, [vp, C.c_size_t, C.c_size_t, C.c_size_t]),
        "ref_zero_grad": (None, [vp]),
        "ref_adam_step": (None, [vp, C.c_int, C.c_int]),
        "ref_upsample_grid": (None, [vp]),
        "ref_upsample_octree": (None, [vp]),
        "ref_rebuild_octree": (None, [vp]),
        "ref_pack_color": (C.c_uint32, [c_fp]),
        "ref_expf": (C.c_float, [C.c_float]),
        "ref_logf": (C.c_float, [C.c_float]),
        "ref_omp_threads": (C.c_int, []),
        "ref_set_omp_threads": (None, [C.c_int]),
        "ref_stbi_loadf": (C.c_int, [C.c_char_p, C.c_int, c_fp, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "ref_save_bmp": (C.c_int, [C.c_char_p, c_u32p, C.c_int, C.c_int]),
    }
    for name, (res, args) in sigs.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _ref[kind] = L
    return L


def np_view(ptr, n, dtype=np.float32):
    """numpy view of n elements at a ctypes pointer (no copy)."""
    ct = {np.float32: C.c_float, np.float64: C.c_double}[dtype]
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,))


def lcg_grid(gs, D=20.0, blob=False, seed=0x9E3779B97F4A7C15, sphere=False):
    """SURVEY 8(d) synthetic grid (numpy restatement, vectorised): [gs^3, 13] float32."""
    n = gs ** 3 * 13
    # LCG is sequential; use the closed-form jump via modular arithmetic on python ints in blocks
    a, c, m = 6364136223846793005, 1442695040888963407, 1 << 64
    out = np.empty(n, dtype=np.uint64)
    s = seed
    # vectorised: s_k = a^k s0 + c (a^k - 1)/(a - 1); do it blockwise with numpy uint64 wraparound
    blk = 1 << 16
    ak = np.empty(blk, dtype=np.uint64)
    ck = np.empty(blk, dtype=np.uint64)
    A, Cc = 1, 0
    for i in range(blk):
        A = (A * a) % m
        Cc = (Cc * a + c) % m
        ak[i] = A
        ck[i] = Cc
    pos = 0
    with np.errstate(over="ignore"):
        while pos < n:
            k = min(blk, n - pos)
            out[pos:pos + k] = ak[:k] * np.uint64(s) + ck[:k]
            s = int(out[pos + k - 1])
            pos += k
    u = ((out >> np.uint64(40)).astype(np.float32) / np.float32(16777216.0)).reshape(-1, 13)
    g = np.empty_like(u)
    g[:, 0] = np.float32(D) * (u[:, 0] - np.float32(0.25))
    g[:, 1:4] = u[:, 1:4]
    g[:, 4] = np.float32(1.0) + np.float32(2.0) * u[:, 4]
    g[:, 5:] = u[:, 5:] - np.float32(0.5)
    if blob:
        idx = np.arange(gs ** 3)
        x, y, z = idx % gs, (idx // gs) % gs, idx // (gs * gs)
        p = [((v.astype(np.float32) + np.float32(0.5)) / np.float32(gs) - np.float32(0.5)) for v in (x, y, z)]
        r2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2]
        g[:, 0] *= np.exp(-(r2) / np.float32(2.0 * 0.2 * 0.2)).astype(np.float32)
    if sphere:   # BASELINE config 4: positive density inside r = 0.288 (volume fraction 0.100), negative outside
        idx = np.arange(gs ** 3)
        x, y, z = idx % gs, (idx // gs) % gs, idx // (gs * gs)
        p = [((v.astype(np.float32) + np.float32(0.5)) / np.float32(gs) - np.float32(0.5)) for v in (x, y, z)]
        r2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2]
        mag = np.abs(g[:, 0]) + np.float32(0.5)
        g[:, 0] = np.where(r2 < np.float32(0.288) * np.float32(0.288), mag, -mag)
    return g


def write_png_rgba8(path, rgba8):
    """minimal PNG writer (8-bit RGBA, filter 0, zlib) so that the reference's own stbi_loadf can be fed seeded images"""
    import struct
    import zlib
    h, w, _ = rgba8.shape
    raw = b"".join(b"\x00" + rgba8[y].tobytes() for y in range(h))

    def chunk(tag, data):
        c = struct.pack(">I", len(data)) + tag + data
        return c + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))
