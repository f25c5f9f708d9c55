"""literf-radiancefields_b200 — Python host side of the B200-native LiteRF ray-march path.

The product is the C-ABI CUDA library (csrc/ -> _build/liblrf_b200.so, declared in include/lrf.h).
This module binds it with ctypes and mirrors the reference's operator interface — class
``RayMarcherExample`` with the method names of example_tracer/example_tracer.h:40-252 — so that
parity tests and the benchmark read like the reference's own main.cpp.  (The C++ drop-in shim for
main.cpp itself is host/example_tracer/.)  There is no CPU path: importing works anywhere, but
constructing a tracer without a B200 raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LRF_B200_LIB") or os.path.join(_HERE, "_build", "liblrf_b200.so")   # the override is for kernel-variant experiments (tools/)

_c_fp = C.POINTER(C.c_float)
_c_u32p = C.POINTER(C.c_uint32)
_c_i32p = C.POINTER(C.c_int32)
_vp = C.c_void_p

# name -> (restype, argtypes); must list every symbol include/lrf.h declares (tests/test_capi_symbols.py checks)
SIGNATURES = {
    "lrf_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "lrf_destroy": (C.c_int, [_vp]),
    "lrf_last_error": (C.c_char_p, []),
    "lrf_sync": (C.c_int, [_vp]),
    "lrf_init_grid": (C.c_int, [_vp, C.c_int]),
    "lrf_init_grad": (C.c_int, [_vp]),
    "lrf_zero_grad": (C.c_int, [_vp]),
    "lrf_set_bbox": (C.c_int, [_vp, _c_fp, _c_fp]),
    "lrf_grid_size": (C.c_int, [_vp]),
    "lrf_grid_upload": (C.c_int, [_vp, _c_fp]),
    "lrf_grid_download": (C.c_int, [_vp, _c_fp]),
    "lrf_grad_upload": (C.c_int, [_vp, _c_fp]),
    "lrf_grad_download": (C.c_int, [_vp, _c_fp]),
    "lrf_adam_state_upload": (C.c_int, [_vp, C.c_int, _c_fp]),
    "lrf_adam_state_download": (C.c_int, [_vp, C.c_int, _c_fp]),
    "lrf_set_boxes": (C.c_int, [_vp, _c_i32p, C.c_size_t]),
    "lrf_set_boxes_dense": (C.c_int, [_vp]),
    "lrf_box_count": (C.c_size_t, [_vp]),
    "lrf_get_boxes": (C.c_int, [_vp, _c_i32p]),
    "lrf_set_view_inv": (C.c_int, [_vp, _c_fp]),
    "lrf_set_proj_inv": (C.c_int, [_vp, _c_fp]),
    "lrf_render": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, C.c_int]),
    "lrf_render_dev_async": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, C.c_int]),
    "lrf_render_views_dev_async": (C.c_int, [_vp, C.c_uint32, _vp, _vp, C.c_uint32, C.c_uint32, C.c_int]),
    "lrf_render_views": (C.c_int, [_vp, C.c_uint32, _vp, _vp, C.c_uint32, C.c_uint32, C.c_int]),
    "lrf_render_rows_dev_async": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]),
    "lrf_render_staged_debug": (C.c_int, [_vp, _vp, _vp, C.c_uint32, C.c_uint32]),
    "lrf_render_lean_debug": (C.c_int, [_vp, _vp, _vp, _vp, C.c_uint32, C.c_uint32]),
    "lrf_render_stats": (C.c_int, [_vp, _vp, C.c_int]),
    "lrf_params_touched": (C.c_int, [_vp]),
    "lrf_render_debug": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_int, _c_u32p, _c_u32p, _c_fp, _c_i32p, C.c_uint32]),
    "lrf_trace_rays": (C.c_int, [_vp, C.c_size_t, _c_fp, _c_fp, C.c_int, _c_u32p, _c_u32p, _c_fp, _c_i32p, C.c_uint32]),
    "lrf_render_grad": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp, C.c_int]),
    "lrf_render_grad_dev_async": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp, C.c_int]),
    "lrf_rays_grad_dev_async": (C.c_int, [_vp, C.c_size_t, _vp, _vp, _vp, C.c_int]),
    "lrf_rays_grad": (C.c_int, [_vp, C.c_size_t, _c_fp, _c_fp, _c_fp, C.c_int]),
    "lrf_gen_rays_dev_async": (C.c_int, [_vp, C.c_size_t, _vp, _vp, C.c_uint32, C.c_uint32, _vp, _vp]),
    "lrf_tv_grad": (C.c_int, [_vp]),
    "lrf_adam_init": (C.c_int, [_vp]),
    "lrf_adam_step": (C.c_int, [_vp, C.c_int, C.c_int]),
    "lrf_train_percell_async": (C.c_int, [_vp, C.c_int]),
    "lrf_set_adam_hparams": (C.c_int, [_vp, C.c_float, C.c_float, C.c_float, C.c_float]),
    "lrf_upsample_grid": (C.c_int, [_vp]),
    "lrf_upsample_boxes": (C.c_int, [_vp]),
    "lrf_rebuild_boxes": (C.c_int, [_vp]),
    "lrf_last_time_ms": (C.c_int, [_vp, C.c_char_p, _c_fp]),
    "lrf_timer_start": (C.c_int, [_vp]),
    "lrf_timer_stop_ms": (C.c_int, [_vp, _c_fp]),
    "lrf_param_device_ptr": (_vp, [_vp]),
    "lrf_grad_device_ptr": (_vp, [_vp]),
    "lrf_param_count": (C.c_size_t, [_vp]),
    "lrf_stream": (_vp, [_vp]),
    "lrf_use_stream": (C.c_int, [_vp, _vp]),
    "lrf_device_alloc": (C.c_int, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "lrf_device_free": (C.c_int, [_vp, _vp]),
    "lrf_host_alloc_pinned": (C.c_int, [C.c_size_t, C.POINTER(_vp)]),
    "lrf_host_free_pinned": (C.c_int, [_vp]),
    "lrf_memcpy_h2d_async": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "lrf_memcpy_d2h_async": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "lrf_launch_count": (C.c_uint64, [_vp]),
    "lrf_set_option": (C.c_int, [_vp, C.c_char_p, C.c_int]),
    "lrf_count_touched_nodes": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_uint64)]),
    "lrf_fill_synthetic": (C.c_int, [_vp, C.c_float, C.c_int]),
    "lrf_rays_grad_tv_dev_async": (C.c_int, [_vp, C.c_size_t, _vp, _vp, _vp, C.c_int]),
    "lrf_adam_zero_async": (C.c_int, [_vp, C.c_int]),
    "lrf_ipc_export": (C.c_int, [_vp, C.c_int, _vp]),
    "lrf_peer_attach_ipc": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp]),
    "lrf_peer_attach_local": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(_vp)]),
    "lrf_peer_detach": (C.c_int, [_vp]),
    "lrf_peer_attach_flags_ipc": (C.c_int, [_vp, _vp]),
    "lrf_peer_barrier_async": (C.c_int, [_vp]),
    "lrf_mc_supported": (C.c_int, [C.c_int]),
    "lrf_mc_prepare": (C.c_int, [_vp, C.c_int, C.c_int]),
    "lrf_mc_create": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "lrf_mc_import": (C.c_int, [_vp, C.c_int]),
    "lrf_mc_add_device": (C.c_int, [_vp]),
    "lrf_mc_bind": (C.c_int, [_vp]),
    "lrf_mc_attach_local": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "lrf_mc_attached": (C.c_int, [_vp]),
    "lrf_peer_barrier_timed_out": (C.c_int, [_vp]),
    "lrf_tv_grad_slab_async": (C.c_int, [_vp]),
    "lrf_adam_peer_async": (C.c_int, [_vp, C.c_int]),
    "lrf_train_percell_peer_async": (C.c_int, [_vp, C.c_int]),
    "lrf_zero_grad_async": (C.c_int, [_vp]),
    "lrf_model_save": (C.c_int, [_vp, C.c_char_p]),
    "lrf_model_load": (C.c_int, [_vp, C.c_char_p]),
    "lrf_targets_resize": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_uint32]),
    "lrf_target_ingest_rgba8": (C.c_int, [_vp, C.c_uint32, _vp, C.c_int]),
    "lrf_target_upload_rgbaf": (C.c_int, [_vp, C.c_uint32, _c_fp]),
    "lrf_target_download": (C.c_int, [_vp, C.c_uint32, _c_fp]),
    "lrf_target_device_ptr": (_vp, [_vp, C.c_uint32]),
    "lrf_render_grad_target_async": (C.c_int, [_vp, C.c_uint32, C.c_int]),
    "lrf_gather_targets_dev_async": (C.c_int, [_vp, C.c_size_t, _vp, _vp]),
    "lrf_save_bmp": (C.c_int, [C.c_char_p, _c_u32p, C.c_int, C.c_int]),
    "lrf_render_bmp": (C.c_int, [_vp, C.c_char_p, C.c_uint32, C.c_uint32, C.c_int]),
}

_lib = None


class LrfError(RuntimeError):
    pass


def load_library(path=None):
    """dlopen liblrf_b200.so and set ctypes signatures.  Fails loudly if the extension was not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise LrfError("liblrf_b200.so is not built (%s): run `python -c 'import __graft_entry__ as g; g.build()'` — there is no CPU fallback" % path)
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        f = getattr(lib, name)
        f.restype = res
        f.argtypes = args
    _lib = lib
    return lib


def _fp(a):
    return a.ctypes.data_as(_c_fp)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# ---- caller-side camera helpers (main.cpp:78,99-102), bit-identical to LiteMath: a small host-only helper library of the harness
# (host/harness/lrf_camera.cpp -> _build/liblrf_camera.so), not part of the product library ----
_cam = None


def _camera_lib():
    global _cam
    if _cam is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "liblrf_camera.so")
        if not os.path.exists(path):
            raise LrfError("liblrf_camera.so is not built (%s): run `python -c 'import __graft_entry__ as g; g.build()'`" % path)
        _cam = C.CDLL(path)
        _cam.lrf_cam_perspective.argtypes = [C.c_float] * 4 + [_c_fp]
        _cam.lrf_cam_orbit_view.argtypes = [C.c_int, C.c_int, _c_fp]
        _cam.lrf_mat4_inverse.argtypes = [_c_fp, _c_fp]
        _cam.lrf_mat4_mul.argtypes = [_c_fp, _c_fp, _c_fp]
        for f in (_cam.lrf_cam_perspective, _cam.lrf_cam_orbit_view, _cam.lrf_mat4_inverse, _cam.lrf_mat4_mul):
            f.restype = None
    return _cam


def perspective(fovy, aspect, zn, zf):
    m = np.zeros(16, np.float32)
    _camera_lib().lrf_cam_perspective(fovy, aspect, zn, zf, _fp(m))
    return m


def orbit_view(k, n_views):
    m = np.zeros(16, np.float32)
    _camera_lib().lrf_cam_orbit_view(k, n_views, _fp(m))
    return m


def inverse4x4(m):
    m = _f32(m).ravel()
    out = np.zeros(16, np.float32)
    _camera_lib().lrf_mat4_inverse(_fp(m), _fp(out))
    return out


def mat4_mul(a, b):
    a, b = _f32(a).ravel(), _f32(b).ravel()
    out = np.zeros(16, np.float32)
    _camera_lib().lrf_mat4_mul(_fp(a), _fp(b), _fp(out))
    return out


class PinnedBuffer:
    """Page-locked host memory viewed as a numpy array (for the end-to-end benchmark)."""

    def __init__(self, shape, dtype):
        self.lib = load_library()
        self.dtype = np.dtype(dtype)
        self.shape = tuple(np.atleast_1d(shape))
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = _vp()
        if self.lib.lrf_host_alloc_pinned(max(self.nbytes, 16), C.byref(p)):
            raise LrfError(self.lib.lrf_last_error().decode())
        self.ptr = p.value
        buf = (C.c_char * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype).reshape(self.shape)

    def free(self):
        if self.ptr:
            self.array = None
            self.lib.lrf_host_free_pinned(self.ptr)
            self.ptr = None


class DeviceBuffer:
    def __init__(self, tracer, nbytes):
        self.t = tracer
        self.nbytes = nbytes
        p = _vp()
        tracer._ck(tracer.lib.lrf_device_alloc(tracer.h, max(nbytes, 16), C.byref(p)))
        self.ptr = p.value

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        self.t._ck(self.t.lib.lrf_memcpy_h2d_async(self.t.h, self.ptr, arr.ctypes.data, arr.nbytes))
        self.t.sync()

    def download(self, arr):
        assert arr.flags.c_contiguous and arr.nbytes <= self.nbytes
        self.t._ck(self.t.lib.lrf_memcpy_d2h_async(self.t.h, arr.ctypes.data, self.ptr, arr.nbytes))
        self.t.sync()
        return arr

    def free(self):
        if self.ptr:
            self.t.lib.lrf_device_free(self.t.h, self.ptr)
            self.ptr = None


class RayMarcherExample:
    """Mirror of the reference class (example_tracer/example_tracer.h:40-252): same method names and
    argument meaning; state lives in HBM.  Matrices are 16 float32, column-major (LiteMath float4x4)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = _vp()
        if self.lib.lrf_create(device, C.byref(h)):
            raise LrfError(self.lib.lrf_last_error().decode())
        self.h = h
        self.device = device
        # reference constructor defaults (example_tracer.h:43-49) are overwritten by every caller; keep identity

    def _ck(self, rc):
        if rc:
            raise LrfError(self.lib.lrf_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.lrf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- hooks of the kernel_slicer convention (example_tracer.h:223-230) --
    def CommitDeviceData(self):
        pass

    def UpdateMembersPlainData(self):
        self.sync()

    def GetExecutionTime(self, name, out=None):
        o = np.zeros(4, np.float32)
        self._ck(self.lib.lrf_last_time_ms(self.h, name.encode(), _fp(o)))
        if out is not None:
            out[:] = o
        return float(o[0])

    def sync(self):
        self._ck(self.lib.lrf_sync(self.h))

    # -- grid --
    def InitGrid(self, gs):
        self._ck(self.lib.lrf_init_grid(self.h, int(gs)))

    def InitGrad(self):
        self._ck(self.lib.lrf_init_grad(self.h))

    def zeroGrad(self):
        self._ck(self.lib.lrf_zero_grad(self.h))

    def SetBoundingBox(self, mn, mx):
        mn, mx = _f32(mn), _f32(mx)
        self._ck(self.lib.lrf_set_bbox(self.h, _fp(mn), _fp(mx)))

    @property
    def gridSize(self):
        return self.lib.lrf_grid_size(self.h)

    def _cells(self):
        return np.empty((self.gridSize ** 3, 13), np.float32)

    def set_grid(self, cells):  # public member `grid`
        cells = _f32(cells)
        assert cells.size == self.gridSize ** 3 * 13
        self._ck(self.lib.lrf_grid_upload(self.h, _fp(cells)))

    def get_grid(self):
        out = self._cells()
        self._ck(self.lib.lrf_grid_download(self.h, _fp(out)))
        return out

    def set_grad(self, cells):  # public member `grid_d`
        cells = _f32(cells)
        self._ck(self.lib.lrf_grad_upload(self.h, _fp(cells)))

    def get_grad(self):
        out = self._cells()
        self._ck(self.lib.lrf_grad_download(self.h, _fp(out)))
        return out

    def set_adam_state(self, V, S):
        V, S = _f32(V), _f32(S)
        self._ck(self.lib.lrf_adam_state_upload(self.h, 0, _fp(V)))
        self._ck(self.lib.lrf_adam_state_upload(self.h, 1, _fp(S)))

    def get_adam_state(self):
        V, S = self._cells(), self._cells()
        self._ck(self.lib.lrf_adam_state_download(self.h, 0, _fp(V)))
        self._ck(self.lib.lrf_adam_state_download(self.h, 1, _fp(S)))
        return V, S

    # -- boxes (public member `boxes`) --
    def set_boxes(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.int32).reshape(-1, 3)
        self._ck(self.lib.lrf_set_boxes(self.h, xyz.ctypes.data_as(_c_i32p), xyz.shape[0]))

    def set_boxes_dense(self):
        self._ck(self.lib.lrf_set_boxes_dense(self.h))

    def box_count(self):
        return self.lib.lrf_box_count(self.h)

    def get_boxes(self):
        out = np.empty((self.box_count(), 3), np.int32)
        self._ck(self.lib.lrf_get_boxes(self.h, out.ctypes.data_as(_c_i32p)))
        return out

    # -- camera: the reference setters take the forward matrix and store its inverse (example_tracer.h:214-215) --
    def SetWorldViewMatrix(self, m):
        inv = inverse4x4(m)
        self._ck(self.lib.lrf_set_view_inv(self.h, _fp(inv)))

    def SetWorldViewMProjatrix(self, m):  # [sic]
        inv = inverse4x4(m)
        self._ck(self.lib.lrf_set_proj_inv(self.h, _fp(inv)))

    def set_view_inv(self, inv):
        inv = _f32(inv).ravel()
        self._ck(self.lib.lrf_set_view_inv(self.h, _fp(inv)))

    def set_proj_inv(self, inv):
        inv = _f32(inv).ravel()
        self._ck(self.lib.lrf_set_proj_inv(self.h, _fp(inv)))

    # -- forward --
    def kernel2D_RayMarch(self, out_color, width, height, greyscale=False):
        """out_color: uint32 numpy array [height*width] (host), or PinnedBuffer.array"""
        assert out_color.dtype == np.uint32 and out_color.size >= width * height and out_color.flags.c_contiguous
        self._ck(self.lib.lrf_render(self.h, out_color.ctypes.data, width, height, int(greyscale)))
        return out_color

    def render_dev_async(self, out_dev_ptr, width, height, greyscale=False):
        self._ck(self.lib.lrf_render_dev_async(self.h, out_dev_ptr, width, height, int(greyscale)))

    def render_views_dev_async(self, view_invs, out_dev_ptr, width, height, greyscale=False):
        """len(view_invs) frames (width*height pixels apart in the device buffer) in one call; view_invs: inverse view matrices, 16 floats each"""
        m = np.ascontiguousarray(np.asarray(view_invs, dtype=np.float32).reshape(-1, 16))
        self._ck(self.lib.lrf_render_views_dev_async(self.h, m.shape[0], m.ctypes.data_as(_vp), out_dev_ptr, width, height, int(greyscale)))

    def render_views(self, view_invs, out_color, width, height, greyscale=False):
        """len(view_invs) frames into the host array out_color (uint32, frame v at v*width*height; PinnedBuffer.array for overlapped copies), synchronous"""
        m = np.ascontiguousarray(np.asarray(view_invs, dtype=np.float32).reshape(-1, 16))
        assert out_color.dtype == np.uint32 and out_color.size >= m.shape[0] * width * height and out_color.flags.c_contiguous
        self._ck(self.lib.lrf_render_views(self.h, m.shape[0], m.ctypes.data_as(_vp), out_color.ctypes.data, width, height, int(greyscale)))
        return out_color

    def render_rows_dev_async(self, out_dev_ptr, width, height, y0, y1, greyscale=False):
        """rows [y0, y1) of the frame into the full-frame device buffer (latency mode: row bands over GPUs)"""
        self._ck(self.lib.lrf_render_rows_dev_async(self.h, out_dev_ptr, width, height, y0, y1, int(greyscale)))

    def render_staged_debug(self, width, height):
        """frame + per-pixel shaded-sample count of the TMA-staged kernel (0xffffffff: pixel rendered by the exact kernel)"""
        out = np.zeros(width * height, np.uint32)
        ns = np.zeros(width * height, np.uint32)
        self._ck(self.lib.lrf_render_staged_debug(self.h, out.ctypes.data, ns.ctypes.data, width, height))
        return out, ns

    def render_lean_debug(self, width, height):
        """frame, per-pixel shaded-sample count (0xffffffff: generic fallback) and final transmittance of the register-window march"""
        out = np.zeros(width * height, np.uint32)
        ns = np.zeros(width * height, np.uint32)
        thr = np.zeros(width * height, np.float32)
        self._ck(self.lib.lrf_render_lean_debug(self.h, out.ctypes.data, ns.ctypes.data, thr.ctypes.data, width, height))
        return out, ns, thr

    def render_stats(self, reset=False):
        """[ambiguous-break rays, other rays sent to the exact kernel, tiles that did not fit the staged box, last render staged?]"""
        st = np.zeros(4, np.uint32)
        self._ck(self.lib.lrf_render_stats(self.h, st.ctypes.data, int(reset)))
        return st

    def params_touched(self):
        self._ck(self.lib.lrf_params_touched(self.h))

    def render_debug(self, width, height, greyscale=False, keep=0):
        n = width * height
        n_hit, n_shaded = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        rgbtc = np.zeros((n, 5), np.float32)
        vox = np.zeros((n, keep, 3), np.int32) if keep else None
        self._ck(self.lib.lrf_render_debug(self.h, width, height, int(greyscale), n_hit.ctypes.data_as(_c_u32p), n_shaded.ctypes.data_as(_c_u32p),
                                           _fp(rgbtc), vox.ctypes.data_as(_c_i32p) if keep else None, keep))
        return n_hit, n_shaded, rgbtc, vox

    def trace_rays(self, o, d, greyscale=False, keep=0):
        o, d = _f32(o).reshape(-1, 3), _f32(d).reshape(-1, 3)
        n = o.shape[0]
        n_hit, n_shaded = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        rgbtc = np.zeros((n, 5), np.float32)
        vox = np.zeros((n, keep, 3), np.int32) if keep else None
        self._ck(self.lib.lrf_trace_rays(self.h, n, _fp(o), _fp(d), int(greyscale), n_hit.ctypes.data_as(_c_u32p), n_shaded.ctypes.data_as(_c_u32p),
                                         _fp(rgbtc), vox.ctypes.data_as(_c_i32p) if keep else None, keep))
        return n_hit, n_shaded, rgbtc, vox

    def count_touched_nodes(self, width, height, greyscale=False):
        out = C.c_uint64(0)
        self._ck(self.lib.lrf_count_touched_nodes(self.h, width, height, int(greyscale), C.byref(out)))
        return out.value

    # -- backward --
    def kernel2D_RayMarchGrad(self, width, height, res, greyscale=False):
        """res: float32 [height*width*4] reference image (float4 per pixel, stbi layout)"""
        res = _f32(res)
        assert res.size >= width * height * 4
        self._ck(self.lib.lrf_render_grad(self.h, width, height, res.ctypes.data, int(greyscale)))

    def rays_grad(self, o, d, ref_rgba, greyscale=False):
        o, d, r = _f32(o).reshape(-1, 3), _f32(d).reshape(-1, 3), _f32(ref_rgba).reshape(-1, 4)
        self._ck(self.lib.lrf_rays_grad(self.h, o.shape[0], _fp(o), _fp(d), _fp(r), int(greyscale)))

    def kernel2D_TVGrad(self):
        self._ck(self.lib.lrf_tv_grad(self.h))

    # -- optimiser --
    def optimizerInit(self):
        self._ck(self.lib.lrf_adam_init(self.h))

    def optimizerStep(self, it):
        self._ck(self.lib.lrf_adam_step(self.h, it, 0))

    def optimizerStepDensity(self, it):
        self._ck(self.lib.lrf_adam_step(self.h, it, 1))

    def optimizerStepSH(self, it):
        self._ck(self.lib.lrf_adam_step(self.h, it, 2))

    # -- level transition --
    def UpsampleGrid(self):
        self._ck(self.lib.lrf_upsample_grid(self.h))

    def UpsampleOctree(self):
        self._ck(self.lib.lrf_upsample_boxes(self.h))

    def RebuildOctree(self):
        self._ck(self.lib.lrf_rebuild_boxes(self.h))

    def timer_start(self):
        self._ck(self.lib.lrf_timer_start(self.h))

    def timer_stop_ms(self):
        ms = C.c_float(0)
        self._ck(self.lib.lrf_timer_stop_ms(self.h, C.byref(ms)))
        return ms.value

    def set_option(self, name, value):
        self._ck(self.lib.lrf_set_option(self.h, name.encode(), int(value)))

    def launch_count(self):
        return self.lib.lrf_launch_count(self.h)

    # ---- multi-GPU exchange over peer memory ----
    def ipc_export(self, which):
        buf = C.create_string_buffer(64)
        self._ck(self.lib.lrf_ipc_export(self.h, which, buf))
        return buf.raw

    def peer_attach_ipc(self, world, rank, param_handles, grad_handles):
        """handles: lists of `world` 64-byte strings in rank order"""
        pb, gb = b"".join(param_handles), b"".join(grad_handles)
        assert len(pb) == 64 * world and len(gb) == 64 * world
        self._ck(self.lib.lrf_peer_attach_ipc(self.h, world, rank, pb, gb))

    def peer_attach_local(self, rank, tracers):
        arr = (_vp * len(tracers))(*[t.h for t in tracers])
        self._ck(self.lib.lrf_peer_attach_local(self.h, len(tracers), rank, arr))

    def peer_attach_flags_ipc(self, flag_handles):
        fb = b"".join(flag_handles)
        self._ck(self.lib.lrf_peer_attach_flags_ipc(self.h, fb))

    # ---- the exchange through NVLink multicast objects (include/lrf.h: lrf_mc_*) ----
    def mc_supported(self):
        return bool(self.lib.lrf_mc_supported(self.device))

    def mc_prepare(self, world, rank):
        self._ck(self.lib.lrf_mc_prepare(self.h, world, rank))

    def mc_create(self):
        fd = C.c_int(-1)
        self._ck(self.lib.lrf_mc_create(self.h, C.byref(fd)))
        return fd.value

    def mc_import(self, fd):
        self._ck(self.lib.lrf_mc_import(self.h, fd))

    def mc_add_device(self):
        self._ck(self.lib.lrf_mc_add_device(self.h))

    def mc_bind(self):
        self._ck(self.lib.lrf_mc_bind(self.h))

    def mc_attached(self):
        return bool(self.lib.lrf_mc_attached(self.h))

    @staticmethod
    def mc_attach_local(tracers):
        """contexts of ONE process (one per device) as a multicast group; rank = position in the list"""
        arr = (_vp * len(tracers))(*[t.h for t in tracers])
        tracers[0]._ck(tracers[0].lib.lrf_mc_attach_local(len(tracers), arr))

    def peer_barrier_async(self):
        self._ck(self.lib.lrf_peer_barrier_async(self.h))

    def peer_barrier_timed_out(self):
        return bool(self.lib.lrf_peer_barrier_timed_out(self.h))

    def peer_detach(self):
        self._ck(self.lib.lrf_peer_detach(self.h))

    def tv_grad_slab_async(self):
        self._ck(self.lib.lrf_tv_grad_slab_async(self.h))

    def adam_peer_async(self, it):
        self._ck(self.lib.lrf_adam_peer_async(self.h, it))

    def train_percell_peer_async(self, it):
        self._ck(self.lib.lrf_train_percell_peer_async(self.h, it))

    def zero_grad_async(self):
        self._ck(self.lib.lrf_zero_grad_async(self.h))

    def rays_grad_tv_dev_async(self, n, o_dev, d_dev, ref_dev, greyscale=False):
        self._ck(self.lib.lrf_rays_grad_tv_dev_async(self.h, n, o_dev, d_dev, ref_dev, int(greyscale)))

    def adam_zero_async(self, it):
        self._ck(self.lib.lrf_adam_zero_async(self.h, it))

    def fill_synthetic(self, density_scale=20.0, mode=0):
        """seeded synthetic grid of SURVEY 8(d) generated on the device (== tests/libs.py::lcg_grid bit for bit)"""
        self._ck(self.lib.lrf_fill_synthetic(self.h, float(density_scale), int(mode)))

    # ---- callers either side of the path ----
    def save_model(self, path):
        """ref: main.cpp:139-141 (raw gs^3 x Cell dump)"""
        self._ck(self.lib.lrf_model_save(self.h, os.fsencode(path)))

    def load_model(self, path):
        """ref: main.cpp:54-56"""
        self._ck(self.lib.lrf_model_load(self.h, os.fsencode(path)))

    def targets_resize(self, n_views, width, height):
        self._ck(self.lib.lrf_targets_resize(self.h, n_views, width, height))

    def target_ingest_rgba8(self, view, rgba8, flip_vertical=True):
        """decoded 8-bit RGBA rows in file order -> resident float4 target (stbi_loadf semantics, main.cpp:72,96)"""
        a = np.ascontiguousarray(rgba8, dtype=np.uint8)
        self._ck(self.lib.lrf_target_ingest_rgba8(self.h, view, a.ctypes.data, int(flip_vertical)))

    def target_upload(self, view, rgba):
        self._ck(self.lib.lrf_target_upload_rgbaf(self.h, view, _fp(_f32(rgba))))

    def target_download(self, view, width, height):
        out = np.zeros((width * height, 4), np.float32)
        self._ck(self.lib.lrf_target_download(self.h, view, _fp(out)))
        return out

    def render_grad_target_async(self, view, greyscale=False):
        self._ck(self.lib.lrf_render_grad_target_async(self.h, view, int(greyscale)))

    def gather_targets_dev_async(self, n, picks_dev_ptr, ref_dev_ptr):
        self._ck(self.lib.lrf_gather_targets_dev_async(self.h, n, picks_dev_ptr, ref_dev_ptr))

    def render_bmp(self, path, width, height, greyscale=False):
        self._ck(self.lib.lrf_render_bmp(self.h, os.fsencode(path), width, height, int(greyscale)))


def save_bmp(path, pixels, width, height):
    """ref: LiteImage::SaveBMP as L1Loss calls it (example_tracer.cpp:614)"""
    a = np.ascontiguousarray(pixels, dtype=np.uint32)
    if load_library().lrf_save_bmp(os.fsencode(path), a.ctypes.data_as(_c_u32p), width, height):
        raise LrfError(load_library().lrf_last_error().decode())


def L1Loss(ref, gen, width, height, pImpl, fileName=None, greyscale=False):
    """ref: example_tracer.cpp:603-625 — backward, then forward render into `gen`.  The reference sets
    *loss = 0 and never computes it (SURVEY §2 defects); returns 0.0 likewise.  BMP output is the caller's
    business here (out of the hot path)."""
    pImpl.kernel2D_RayMarchGrad(width, height, ref, greyscale)
    pImpl.kernel2D_RayMarch(gen, width, height, greyscale)
    return 0.0
