"""ctypes binding of libgrrt.so -- the Python mirror of the reference's host-side interface.

The reference is a Rust binary; its public surface for this path is the CLI (src/cli/cli.rs) and
`Raytracer::{render_section, integrate_ray_at_point}` (src/rendering/raytracer.rs:460-507).  This
module exposes the same operations over the C ABI of include/grrt.h.  It never falls back to a CPU
implementation: if the CUDA library cannot be loaded or no device is present the calls raise.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class GrrtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"grrt error {code}: {msg}")
        self.code = code


def lib_path():
    # GRRT_LIB: developer knob, points at an alternative build of the same library (tools/variants.py)
    return os.environ.get("GRRT_LIB") or os.path.join(_HERE, "libgrrt.so")


def lib():
    """Load libgrrt.so (building it when missing and nvcc is available).  Fails loudly otherwise."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        from . import build as _build
        _build.build()
    L = C.CDLL(path)
    vp, dp, u8p, u32p, u64p = C.c_void_p, abi.c_double_p, abi.c_uint8_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    scene_p = C.POINTER(abi.GrrtScene)
    L.grrt_version.restype = C.c_char_p
    L.grrt_init.argtypes = [C.c_int, C.POINTER(vp)]
    L.grrt_destroy.argtypes = [vp]
    L.grrt_set_option.argtypes = [vp, C.c_char_p, C.c_int]
    L.grrt_kernel_launch_count.argtypes = [vp]
    L.grrt_kernel_launch_count.restype = C.c_ulonglong
    L.grrt_last_error.argtypes = [vp]
    L.grrt_last_error.restype = C.c_char_p
    L.grrt_render.argtypes = [vp, scene_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, dp, u8p, u8p, u32p,
                              C.POINTER(abi.GrrtStats)]
    L.grrt_tonemap.argtypes = [vp, dp, C.c_size_t, C.c_int, C.c_double, u8p]
    L.grrt_trace_ray.argtypes = [vp, scene_p, dp, dp, C.POINTER(abi.GrrtStep), C.c_uint64, u64p, C.POINTER(C.c_int32)]
    L.grrt_trace_pixel_ray.argtypes = [vp, scene_p, C.c_int64, C.c_int64, C.POINTER(abi.GrrtStep), C.c_uint64, u64p,
                                       C.POINTER(C.c_int32)]
    L.grrt_scene_upload.argtypes = [vp, scene_p, C.POINTER(vp)]
    L.grrt_scene_free.argtypes = [vp, vp]
    L.grrt_scene_set_camera.argtypes = [vp, vp, C.POINTER(abi.GrrtCamera)]
    L.grrt_render_resident.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, dp, u8p, u8p, u32p,
                                       C.POINTER(abi.GrrtStats)]
    L.grrt_trace_items_device.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                          C.c_uint32, C.c_uint32, vp, vp, vp, vp, vp, vp]
    L.grrt_trace_items_device_peers.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                                C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp),
                                                C.POINTER(vp), vp, vp]
    L.grrt_order_by_cost_device.argtypes = [vp, vp, C.c_uint32, vp, vp, vp]
    L.grrt_init_multi.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(vp)]
    L.grrt_device_count.argtypes = [vp]
    L.grrt_ipc_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp), C.c_char_p]
    L.grrt_ipc_open.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    L.grrt_ipc_close.argtypes = [vp, vp]
    L.grrt_ipc_free.argtypes = [vp, vp]
    L.grrt_select_mask_device.argtypes = [vp, vp, vp, vp, C.c_uint32, C.c_uint32, vp, vp, dp, vp]
    L.grrt_tonemap_device.argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_double, vp, vp]
    L.grrt_measure_fp64_peak.argtypes = [vp, dp, dp]
    L.grrt_debug_pow.argtypes = [vp, dp, dp, dp, C.c_size_t]
    L.grrt_debug_sincos.argtypes = [vp, dp, dp, dp, C.c_size_t]
    for name in ("grrt_cli_opts_default", "grrt_host_scene_from_toml_file", "grrt_host_scene_from_toml_string",
                 "grrt_host_scene_get", "grrt_host_scene_free", "grrt_host_ray_at", "grrt_host_write_ray_csv",
                 "grrt_host_write_png_rgb8", "grrt_host_write_hdr_xyz", "grrt_host_read_png_rgba8", "grrt_host_free"):
        if not hasattr(L, name):
            continue
    if hasattr(L, "grrt_host_scene_from_toml_file"):
        L.grrt_cli_opts_default.argtypes = [C.POINTER(abi.GrrtCliOpts)]
        L.grrt_host_scene_from_toml_file.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(abi.GrrtCliOpts), C.POINTER(vp),
                                                     C.c_char_p, C.c_size_t]
        L.grrt_host_scene_from_toml_string.argtypes = L.grrt_host_scene_from_toml_file.argtypes
        L.grrt_host_scene_get.argtypes = [vp]
        L.grrt_host_scene_get.restype = scene_p
        L.grrt_host_scene_free.argtypes = [vp]
        L.grrt_host_scene_set_camera.argtypes = [vp, C.POINTER(abi.GrrtCliOpts), C.c_char_p, C.c_size_t]
        L.grrt_host_blackbody_xyz.argtypes = [C.c_double, C.c_double, dp]
        L.grrt_host_blackbody_xyz.restype = None
        L.grrt_host_xyz_to_srgb8.argtypes = [dp, C.c_double, u8p]
        L.grrt_host_xyz_to_srgb8.restype = None
        L.grrt_host_ray_at.argtypes = [C.POINTER(abi.GrrtGeometry), dp, dp, dp, dp, C.c_char_p, C.c_size_t]
        L.grrt_host_write_ray_csv.argtypes = [C.POINTER(abi.GrrtGeometry), C.POINTER(abi.GrrtStep), C.c_uint64, C.c_char_p]
        L.grrt_host_write_png_rgb8.argtypes = [C.c_char_p, u8p, C.c_uint32, C.c_uint32]
        L.grrt_host_write_hdr_xyz.argtypes = [C.c_char_p, dp, C.c_uint32, C.c_uint32]
        L.grrt_host_read_png_rgba8.argtypes = [C.c_char_p, C.POINTER(u8p), u32p, u32p]
        L.grrt_host_free.argtypes = [vp]
    _LIB = L
    return L


def _scene_ptr(scene):
    """Accepts a HostScene, a ctypes GrrtScene, or any object with a `.c` GrrtScene (test fixtures)."""
    if isinstance(scene, HostScene):
        return scene.ptr
    if isinstance(scene, abi.GrrtScene):
        return C.byref(scene)
    if hasattr(scene, "c"):
        return C.byref(scene.c)
    return scene


def _scene_struct(scene):
    if isinstance(scene, HostScene):
        return scene.ptr.contents
    if isinstance(scene, abi.GrrtScene):
        return scene
    return scene.c


class HostScene:
    """create_scene (reference src/cli/shared.rs:131-321) done by the C++ host code of libgrrt."""

    def __init__(self, toml, opts=None, resource_root=None, is_text=False):
        L = lib()
        if not hasattr(L, "grrt_host_scene_from_toml_file"):
            raise GrrtError(abi.ERR_INVALID_ARGUMENT, "libgrrt.so was built without the host scene code")
        self.opts = opts or abi.default_cli_opts()
        self._h = C.c_void_p()
        err = C.create_string_buffer(1024)
        root = None if resource_root is None else os.fsencode(resource_root)
        fn = L.grrt_host_scene_from_toml_string if is_text else L.grrt_host_scene_from_toml_file
        src = toml.encode() if is_text else os.fsencode(toml)
        rc = fn(src, root, C.byref(self.opts), C.byref(self._h), err, len(err))
        if rc:
            raise GrrtError(rc, err.value.decode(errors="replace"))
        self.ptr = L.grrt_host_scene_get(self._h)

    @property
    def c(self):
        return self.ptr.contents

    def set_camera(self, **overrides):
        """New camera for the assembled scene (camera_position, phi, theta, psi, width, height of the CLI options);
        textures and look-up tables are kept.  One call per animation frame."""
        for k, v in overrides.items():
            if k == "camera_position":
                for i in range(3):
                    self.opts.camera_position[i] = float(v[i])
            else:
                setattr(self.opts, k, v)
        err = C.create_string_buffer(1024)
        rc = lib().grrt_host_scene_set_camera(self._h, C.byref(self.opts), err, len(err))
        if rc:
            raise GrrtError(rc, err.value.decode(errors="replace"))

    def close(self):
        if getattr(self, "_h", None):
            lib().grrt_host_scene_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One GrrtCtx bound to one CUDA device."""

    def __init__(self, device=0):
        """`device`: a CUDA device ordinal, or a sequence of ordinals for one multi-device context (grrt_init_multi):
        render() then shards every frame over those GPUs inside the library."""
        self._h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            ids = (C.c_int * len(device))(*[int(d) for d in device])
            rc = lib().grrt_init_multi(len(device), ids, C.byref(self._h))
            self.device = int(device[0]) if len(device) else 0
        else:
            rc = lib().grrt_init(device, C.byref(self._h))
            self.device = device
        if rc:
            raise GrrtError(rc, "grrt_init failed: no usable sm_100 CUDA device (this library has no CPU fallback)")

    @property
    def device_count(self):
        return int(lib().grrt_device_count(self._h))

    def _check(self, rc):
        if rc < 0:
            raise GrrtError(rc, lib().grrt_last_error(self._h).decode(errors="replace"))
        return rc

    def close(self):
        if getattr(self, "_h", None):
            lib().grrt_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- Raytracer::render_section_to_cie_buffer -------------------------------------------------
    def render(self, scene, from_row=None, from_col=None, to_row=None, to_col=None, out=None):
        """`out` (optional): dict of preallocated C-contiguous host arrays xyza[n,4] f64, ray_class[n] u8, stop[n] u8,
        steps[n] u32 -- e.g. views of pinned torch tensors -- to avoid per-call allocation and pageable copies."""
        sc = _scene_struct(scene)
        call = lambda *a: lib().grrt_render(self._h, _scene_ptr(scene), *a)
        return self._render(call, sc.camera.rows, sc.camera.cols, from_row, from_col, to_row, to_col, out)

    def render_resident(self, dscene, rows, cols, from_row=None, from_col=None, to_row=None, to_col=None, out=None):
        """grrt_render_resident: the frame of an uploaded scene (see upload / set_camera) into host buffers."""
        call = lambda *a: lib().grrt_render_resident(self._h, dscene, *a)
        return self._render(call, rows, cols, from_row, from_col, to_row, to_col, out)

    def set_camera(self, dscene, camera):
        """Replaces the camera of an uploaded scene (`camera`: abi.GrrtCamera, e.g. HostScene.c.camera)."""
        self._check(lib().grrt_scene_set_camera(self._h, dscene, C.byref(camera)))

    def _render(self, call, rows, cols, from_row, from_col, to_row, to_col, out):
        fr = 0 if from_row is None else from_row
        fc = 0 if from_col is None else from_col
        tr = rows if to_row is None else to_row
        tc = cols if to_col is None else to_col
        if tr < fr or tc < fc:  # what grrt_render answers; raised here before any buffer is sized from it
            raise GrrtError(abi.ERR_INVALID_ARGUMENT, "inverted section")
        n = (tr - fr) * (tc - fc)
        if out is not None:
            xyza, cls, stop, steps = out["xyza"], out["ray_class"], out["stop"], out["steps"]
            assert xyza.size == 4 * n and cls.size == n and stop.size == n and steps.size == n
        else:
            xyza = np.zeros((n, 4))
            cls = np.zeros(n, dtype=np.uint8)
            stop = np.zeros(n, dtype=np.uint8)
            steps = np.zeros(n, dtype=np.uint32)
        stats = abi.GrrtStats()
        self._check(call(fr, fc, tr, tc, xyza.ctypes.data_as(abi.c_double_p), cls.ctypes.data_as(abi.c_uint8_p),
                         stop.ctypes.data_as(abi.c_uint8_p), steps.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(stats)))
        shape = (tr - fr, tc - fc)
        return dict(xyza=xyza.reshape(shape + (4,)), ray_class=cls.reshape(shape), stop=stop.reshape(shape),
                    steps=steps.reshape(shape), stats=stats.as_dict())

    def tonemap(self, xyza, method=abi.TONEMAP_REINHARD, exposure=1.0):
        a = np.ascontiguousarray(xyza, dtype=np.float64).reshape(-1, 4)
        rgb = np.zeros((a.shape[0], 3), dtype=np.uint8)
        self._check(lib().grrt_tonemap(self._h, a.ctypes.data_as(abi.c_double_p), a.shape[0], method, exposure,
                                       rgb.ctypes.data_as(abi.c_uint8_p)))
        return rgb.reshape(np.shape(xyza)[:-1] + (3,))

    # -- Integrator::integrate with stored steps (render-ray / render-ray-at) -------------------------
    def trace_ray(self, scene, pos=None, mom=None, pixel=None, cap=None):
        sc = _scene_struct(scene)
        cap = int(cap or sc.integration.max_steps)
        buf = (abi.GrrtStep * cap)()
        n, stop = C.c_uint64(), C.c_int32()
        if pixel is not None:
            err = self._check(lib().grrt_trace_pixel_ray(self._h, _scene_ptr(scene), pixel[0], pixel[1], buf, cap,
                                                         C.byref(n), C.byref(stop)))
        else:
            p = (C.c_double * 4)(*pos)
            m = (C.c_double * 4)(*mom)
            err = self._check(lib().grrt_trace_ray(self._h, _scene_ptr(scene), p, m, buf, cap, C.byref(n), C.byref(stop)))
        m_ = min(n.value, cap)
        dt = np.dtype([("x", "f8", 4), ("p", "f8", 4), ("t", "f8"), ("step", "u8")])
        return np.frombuffer(buf, dtype=dt, count=m_).copy(), stop.value, err

    def set_option(self, name, value):
        self._check(lib().grrt_set_option(self._h, name.encode(), int(value)))

    def launch_count(self):
        return int(lib().grrt_kernel_launch_count(self._h))

    def debug_pow(self, x, y):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(np.broadcast_to(y, x.shape), dtype=np.float64)
        out = np.zeros_like(x)
        self._check(lib().grrt_debug_pow(self._h, x.ctypes.data_as(abi.c_double_p), y.ctypes.data_as(abi.c_double_p),
                                         out.ctypes.data_as(abi.c_double_p), x.size))
        return out

    def debug_sincos(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        sn, cs = np.zeros_like(x), np.zeros_like(x)
        self._check(lib().grrt_debug_sincos(self._h, x.ctypes.data_as(abi.c_double_p), sn.ctypes.data_as(abi.c_double_p),
                                            cs.ctypes.data_as(abi.c_double_p), x.size))
        return sn, cs

    def measure_fp64_peak(self):
        rate, ms = C.c_double(), C.c_double()
        self._check(lib().grrt_measure_fp64_peak(self._h, C.byref(rate), C.byref(ms)))
        return rate.value, ms.value

    # -- device-resident API (bench / multi-GPU) --------------------------------------------------------
    def upload(self, scene):
        h = C.c_void_p()
        self._check(lib().grrt_scene_upload(self._h, _scene_ptr(scene), C.byref(h)))
        return h

    def free_scene(self, h):
        lib().grrt_scene_free(self._h, h)

    def trace_items_device(self, dscene, d_items, first_item, n_items, rect, spa, d_xyza, d_class=None, d_stop=None,
                           d_steps=None, d_counters=None, stream=None):
        fr, fc, tr, tc = rect
        return self._check(lib().grrt_trace_items_device(self._h, dscene, d_items, first_item, n_items, fr, fc, tr, tc,
                                                         spa, d_xyza, d_class, d_stop, d_steps, d_counters, stream))

    def trace_items_device_peers(self, dscene, d_items, n_items, rect, spa, xyza_peers, class_peers, stop_peers,
                                 d_counters=None, stream=None, steps_peers=None, item_offset=0, item_stride=1):
        """grrt_trace_items_device_peers: results go at their pixel id into every peer's frame buffers; item i is
        d_items[item_offset + i * item_stride]."""
        fr, fc, tr, tc = rect
        n = len(xyza_peers)
        arr = lambda ptrs: (C.c_void_p * n)(*[int(p) for p in ptrs])
        return self._check(lib().grrt_trace_items_device_peers(self._h, dscene, d_items, item_offset, item_stride, n_items,
                                                               fr, fc, tr, tc, spa, n,
                                                               arr(xyza_peers), arr(class_peers), arr(stop_peers),
                                                               arr(steps_peers) if steps_peers else None,
                                                               d_counters, stream))

    def order_by_cost_device(self, d_selected, n_selected, d_steps, d_ordered, stream=None):
        """grrt_order_by_cost_device: flagged pixel ids by descending base-pass step count (stable)."""
        return self._check(lib().grrt_order_by_cost_device(self._h, d_selected, n_selected, d_steps, d_ordered, stream))

    # -- frame buffers shared between the processes of one box (cudaIpc) ------------------------------
    def ipc_alloc(self, nbytes):
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        self._check(lib().grrt_ipc_alloc(self._h, nbytes, C.byref(ptr), handle))
        return ptr.value, handle.raw

    def ipc_open(self, handle):
        ptr = C.c_void_p()
        self._check(lib().grrt_ipc_open(self._h, handle, C.byref(ptr)))
        return ptr.value

    def ipc_close(self, ptr):
        self._check(lib().grrt_ipc_close(self._h, C.c_void_p(ptr)))

    def ipc_free(self, ptr):
        self._check(lib().grrt_ipc_free(self._h, C.c_void_p(ptr)))

    def select_mask_device(self, dscene, d_xyza, d_class, rows, cols, d_selected, d_count, stream=None):
        ml = C.c_double()
        self._check(lib().grrt_select_mask_device(self._h, dscene, d_xyza, d_class, rows, cols, d_selected, d_count,
                                                  C.byref(ml), stream))
        return ml.value

    def tonemap_device(self, d_xyza, n, method, exposure, d_rgb, stream=None):
        return self._check(lib().grrt_tonemap_device(self._h, d_xyza, n, method, exposure, d_rgb, stream))
