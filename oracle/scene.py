"""Oracle-side scene assembly (TEST INFRASTRUCTURE ONLY -- see grrt_oracle.cpp header).

Restates, independently of the product's C++ host code, what the reference does between
`toml::from_str::<RenderConfig>` and `Scene::new`:

  src/main.rs:74-115                TOML -> RenderConfig, Cartesian camera position
  src/configuration.rs:4-219        schema + defaults (+ AdaptiveSamplingConfig::validate :61-93)
  src/cli/{schwarzschild,kerr,kerr_bl,euclidean,euclidean_spherical}.rs   chart of the camera
  src/cli/shared.rs:23-86           camera four-velocity, future-directedness
  src/cli/shared.rs:131-321         create_scene: textures, objects, LUTs, Camera::new(PI/4, rows=height, cols=width)

TOML parsing is `tomllib`, PNG decoding is PIL (lossless, so equivalent to the `image`
crate); every floating-point step is done by the oracle library (glibc libm, no FMA).
"""
import ctypes as C
import math
import os
import tomllib

import numpy as np

from gr_raytracer_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OracleError(Exception):
    def __init__(self, code, msg):
        super().__init__(f"{msg} (code {code})")
        self.code = code


def lib():
    """Load (building if needed) oracle/libgrrt_oracle.so."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_HERE, "libgrrt_oracle.so")
    src = os.path.join(_HERE, "grrt_oracle.cpp")
    if (not os.path.exists(path)) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(path)):
        import subprocess
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    L = C.CDLL(path)
    d, dp, u8p = C.c_double, abi.c_double_p, abi.c_uint8_p
    L.orc_num_threads.restype = C.c_int
    L.orc_camera_position_native.argtypes = [C.POINTER(abi.GrrtGeometry), dp, dp]
    L.orc_resolve_camera_velocity.argtypes = [C.POINTER(abi.GrrtGeometry), dp, C.c_int, dp, dp]
    L.orc_camera_new.argtypes = [C.POINTER(abi.GrrtGeometry), dp, dp, d, C.c_int64, C.c_int64, d, d, d,
                                 C.POINTER(abi.GrrtCamera)]
    L.orc_tetrad_at.argtypes = [C.POINTER(abi.GrrtGeometry), dp, dp]
    L.orc_inner_product.argtypes = [C.POINTER(abi.GrrtGeometry), dp, dp, dp]
    L.orc_inner_product.restype = d
    L.orc_camera_ray.argtypes = [C.POINTER(abi.GrrtCamera), C.c_int64, C.c_int64, C.c_int, d, d, dp, dp]
    L.orc_build_blackbody_lut.argtypes = [dp, dp]
    L.orc_blackbody_xyz.argtypes = [d, d, dp]
    L.orc_r_isco.argtypes = [d, d]
    L.orc_r_isco.restype = d
    L.orc_circular_orbit.argtypes = [d, d, d, dp]
    L.orc_build_temperature_lut.argtypes = [d, d, d, d, dp, dp, dp]
    L.orc_compute_temperature.argtypes = [C.POINTER(abi.GrrtTemperature), d, dp]
    L.orc_perlin_permutation.argtypes = [C.c_uint32, u8p]
    L.orc_perlin.argtypes = [u8p, d, d, d]
    L.orc_perlin.restype = d
    L.orc_volumetric_basis.argtypes = [dp, dp, dp, dp]
    L.orc_volumetric_density.argtypes = [C.POINTER(abi.GrrtObject), dp]
    L.orc_volumetric_density.restype = d
    L.orc_ks_metrics.argtypes = [d, d, d, d, d, dp, dp]
    L.orc_axial_killing_vector.argtypes = [C.POINTER(abi.GrrtGeometry), dp, dp]
    L.orc_circular_orbit_velocity.argtypes = [C.POINTER(abi.GrrtGeometry), dp, dp]
    L.orc_object_intersects.argtypes = [C.POINTER(abi.GrrtObject), dp, dp, dp, dp]
    L.orc_volumetric_raymarch.argtypes = [C.POINTER(abi.GrrtScene), C.c_uint32, dp, dp, dp, C.c_int, dp, C.POINTER(C.c_uint64)]
    L.orc_srgb_to_xyz.argtypes = [C.c_uint8, C.c_uint8, C.c_uint8, dp]
    L.orc_blend.argtypes = [dp, dp, dp]
    L.orc_texture_color.argtypes = [C.POINTER(abi.GrrtScene), C.POINTER(abi.GrrtTexture), d, d, d, d, dp]
    L.orc_hash_pixel_samples.argtypes = [C.c_int64, C.c_int64, C.c_uint64]
    L.orc_hash_pixel_samples.restype = d
    L.orc_stratified_sample_offset.argtypes = [C.c_int64, C.c_int64, C.c_uint64, C.c_uint64, C.c_uint64, dp, dp]
    L.orc_should_supersample_pair.argtypes = [dp, C.c_int, dp, C.c_int, C.POINTER(abi.GrrtAdaptive), d]
    L.orc_rkf45_simple_equation.argtypes = [d, d, d, dp]
    L.orc_should_stop.argtypes = [C.POINTER(abi.GrrtGeometry), dp, C.c_uint64, C.c_uint64, d]
    L.orc_geodesic_rhs.argtypes = [C.POINTER(abi.GrrtScene), dp, dp, dp, dp]
    L.orc_solver_constants.argtypes = [C.POINTER(abi.GrrtScene), dp, dp, dp, dp]
    L.orc_trace_ray.argtypes = [C.POINTER(abi.GrrtScene), dp, dp, C.POINTER(abi.GrrtStep), C.c_uint64,
                                C.POINTER(C.c_uint64), C.POINTER(C.c_int32)]
    L.orc_trace_pixel_ray.argtypes = [C.POINTER(abi.GrrtScene), C.c_int64, C.c_int64, C.POINTER(abi.GrrtStep),
                                      C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_int32)]
    L.orc_step_to_cartesian.argtypes = [C.POINTER(abi.GrrtGeometry), dp, dp]
    L.orc_color_of_ray.argtypes = [C.POINTER(abi.GrrtScene), C.c_int64, C.c_int64, C.c_int, d, d, dp,
                                   C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_uint64),
                                   C.POINTER(C.c_uint64)]
    L.orc_render.argtypes = [C.POINTER(abi.GrrtScene), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, dp, u8p,
                             u8p, C.POINTER(C.c_uint32), C.POINTER(abi.GrrtStats), C.c_int]
    L.orc_select_mask.argtypes = [C.POINTER(abi.GrrtAdaptive), dp, u8p, C.c_uint32, C.c_uint32,
                                  C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), dp]
    L.orc_tonemap.argtypes = [dp, C.c_size_t, C.c_int, d, u8p]
    L.orc_xyz_to_srgb.argtypes = [dp, d, u8p]
    _LIB = L
    return L


def dptr(a):
    return a.ctypes.data_as(abi.c_double_p)


def u8ptr(a):
    return a.ctypes.data_as(abi.c_uint8_p)


def darr(vals):
    return (C.c_double * len(vals))(*[float(v) for v in vals])


# ---- geometry / camera ------------------------------------------------------------------

def make_geometry(kind, radius=0.0, a=0.0, horizon_epsilon=0.0):
    g = abi.GrrtGeometry()
    g.kind, g.radius, g.a, g.horizon_epsilon = kind, radius, a, horizon_epsilon
    return g


def geometry_from_config(gt):
    """configuration.rs:110-159 (externally tagged enum: "Euclidean" or {Kerr = {...}})."""
    if isinstance(gt, str):
        name, body = gt, {}
    else:
        (name, body), = gt.items()
    kind = abi.GEOM_NAMES.index(name)
    return make_geometry(kind, body.get("radius", 0.0), body.get("a", 0.0), body.get("horizon_epsilon", 0.0))


def camera_position_native(geom, cart):
    out = (C.c_double * 4)()
    lib().orc_camera_position_native(C.byref(geom), darr(cart), out)
    return list(out)


def resolve_camera_velocity(geom, pos, cfg="StaticObserver"):
    """cli/shared.rs:48-86."""
    mode, comps = abi.CAMVEL_STATIC, [0.0] * 4
    if cfg == "Zamo":
        mode = abi.CAMVEL_ZAMO
    elif isinstance(cfg, dict) and "Explicit" in cfg:
        mode, comps = abi.CAMVEL_EXPLICIT, cfg["Explicit"]["components"]
    elif cfg != "StaticObserver":
        raise OracleError(abi.ERR_TOML, f"unknown camera_velocity {cfg!r}")
    out = (C.c_double * 4)()
    rc = lib().orc_resolve_camera_velocity(C.byref(geom), darr(pos), mode, darr(comps), out)
    if rc:
        raise OracleError(rc, "camera four-velocity invalid (normalisation / time orientation)")
    return list(out)


def make_camera(geom, pos, vel, alpha, rows, cols, phi=0.0, theta=0.0, psi=0.0):
    cam = abi.GrrtCamera()
    rc = lib().orc_camera_new(C.byref(geom), darr(pos), darr(vel), alpha, rows, cols, phi, theta, psi, C.byref(cam))
    if rc:
        raise OracleError(rc, "Tetrad is not orthonormal")
    return cam


# ---- scene ------------------------------------------------------------------------------

class Scene:
    """Owns a GrrtScene plus every buffer it points to."""

    def __init__(self):
        self.c = abi.GrrtScene()
        self._keep = []
        self._bitmaps = {}   # path -> (beaming_exponent, ndarray)  TextureMapperFactory, texture.rs:260-296
        self._bb = None

    @property
    def ptr(self):
        return C.byref(self.c)

    def keep(self, a):
        self._keep.append(a)
        return a

    def blackbody_lut(self):
        if self._bb is None:
            lt = np.zeros(abi.BLACKBODY_LUT_N)
            xyz = np.zeros(abi.BLACKBODY_LUT_N * 3)
            lib().orc_build_blackbody_lut(dptr(lt), dptr(xyz))
            self._bb = (self.keep(lt), self.keep(xyz))
            self.c.bb_lut_n = abi.BLACKBODY_LUT_N
            self.c.bb_lut_log_t = dptr(lt)
            self.c.bb_lut_xyz = dptr(xyz)
        return self._bb

    def texture(self, cfg, resource_root=None):
        """get_texture_mapper, cli/shared.rs:323-350."""
        t = abi.GrrtTexture()
        (kind, body), = cfg.items()
        if kind == "Bitmap":
            path = body["path"]
            if path not in self._bitmaps:
                from PIL import Image
                full = path if (os.path.isabs(path) or resource_root is None) else os.path.join(resource_root, path)
                try:
                    img = np.ascontiguousarray(np.array(Image.open(full).convert("RGBA"), dtype=np.uint8))
                except OSError as e:
                    raise OracleError(abi.ERR_TEXTURE, f"Texture error: {e}")
                self._bitmaps[path] = (float(body["beaming_exponent"]), self.keep(img))
            beaming, img = self._bitmaps[path]
            t.kind = abi.TEX_BITMAP
            t.beaming_exponent = beaming        # first user of a file wins (factory cache)
            t.height, t.width = img.shape[0], img.shape[1]
            t.rgba = u8ptr(img)
        elif kind == "Checker":
            t.kind = abi.TEX_CHECKER
            t.beaming_exponent = body["beaming_exponent"]
            t.checker_width, t.checker_height = body["width"], body["height"]
            for dst, col in ((t.c1, body["color1"]), (t.c2, body["color2"])):
                out = (C.c_double * 4)()
                lib().orc_srgb_to_xyz(col[0], col[1], col[2], out)
                dst[:] = list(out)
        elif kind == "BlackBody":
            t.kind = abi.TEX_BLACKBODY
            t.beaming_exponent = body["beaming_exponent"]
            self.blackbody_lut()
        else:
            raise OracleError(abi.ERR_TOML, f"unknown texture {kind}")
        return t

    def temperature(self, geom, temperature, outer_radius):
        """Geometry::get_temperature_computer (schwarzschild.rs:267-279, kerr.rs:498-510,
        kerr_bl.rs:412-424, euclidean.rs:218-225)."""
        tc = abi.GrrtTemperature()
        if geom.kind in (abi.GEOM_EUCLIDEAN, abi.GEOM_EUCLIDEAN_SPHERICAL):
            tc.kind, tc.constant = abi.TEMP_CONSTANT, temperature
            return tc
        a = 0.0 if geom.kind == abi.GEOM_SCHWARZSCHILD else geom.a
        r = np.zeros(abi.TEMPERATURE_LUT_N)
        t = np.zeros(abi.TEMPERATURE_LUT_N)
        isco = C.c_double()
        rc = lib().orc_build_temperature_lut(temperature, outer_radius, a, geom.radius, dptr(r), dptr(t),
                                             C.byref(isco))
        if rc:
            raise OracleError(rc, "temperature LUT construction failed")
        self.keep(r), self.keep(t)
        tc.kind, tc.n, tc.r_isco = abi.TEMP_LUT, abi.TEMPERATURE_LUT_N, isco.value
        tc.lut_r, tc.lut_t = dptr(r), dptr(t)
        return tc

    def add_object(self, obj):
        if self.c.n_objects >= abi.MAX_OBJECTS:
            raise OracleError(abi.ERR_INVALID_CONFIGURATION, "too many objects")
        self.c.objects[self.c.n_objects] = obj
        self.c.n_objects += 1


def _adaptive_from_config(cfg, opts):
    """configuration.rs:21-93 + cli/shared.rs:138-145."""
    ad = abi.GrrtAdaptive()
    ad.enabled = 1 if cfg.get("enabled", True) else 0
    ad.samples_per_axis = cfg.get("samples_per_axis", 4)
    ad.luminance_contrast_threshold = cfg.get("luminance_contrast_threshold", 0.15)
    ad.opacity_contrast_threshold = cfg.get("opacity_contrast_threshold", 0.1)
    ml = cfg.get("minimum_luminance", None)
    ad.has_minimum_luminance = 0 if ml is None else 1
    ad.minimum_luminance = 0.0 if ml is None else ml
    ad.object_hit_opacity_threshold = cfg.get("object_hit_opacity_threshold", 0.5)
    ad.exclude_background_contrast = 1 if cfg.get("exclude_background_contrast", True) else 0
    if ad.samples_per_axis == 0:
        raise OracleError(abi.ERR_INVALID_CONFIGURATION, "adaptive_sampling.samples_per_axis must be greater than zero")
    for name in ("luminance_contrast_threshold", "opacity_contrast_threshold", "object_hit_opacity_threshold"):
        v = getattr(ad, name)
        if not math.isfinite(v) or not (0.0 <= v <= 1.0):
            raise OracleError(abi.ERR_INVALID_CONFIGURATION, f"adaptive_sampling.{name} must be finite and between 0 and 1")
    if ml is not None and (not math.isfinite(ml) or ml < 0.0):
        raise OracleError(abi.ERR_INVALID_CONFIGURATION, "adaptive_sampling.minimum_luminance must be finite and non-negative")
    ad.show_sampling_mask = 1 if opts.show_sampling_mask else 0
    if opts.show_sampling_mask:
        out = (C.c_double * 4)()
        col = list(opts.sampling_mask_color)
        lib().orc_srgb_to_xyz(col[0], col[1], col[2], out)
        ad.mask_xyza[:] = [out[0], out[1], out[2], 255.0 / 255.0]   # CIETristimulus::from_color
    return ad


def scene_from_config(config, opts, resource_root=None):
    """create_scene, cli/shared.rs:131-321 (config = parsed TOML dict, opts = GrrtCliOpts)."""
    sc = Scene()
    geom = geometry_from_config(config.get("geometry_type", "Euclidean"))
    sc.c.geometry = geom
    sc.c.adaptive = _adaptive_from_config(config.get("adaptive_sampling", {}), opts)
    sc.c.integration.max_steps = opts.max_steps
    sc.c.integration.max_radius = opts.max_radius
    sc.c.integration.step_size = opts.step_size
    sc.c.integration.epsilon = opts.epsilon
    sc.c.kerr_rhs_mode = abi.KERR_RHS_REFERENCE_FD

    sc.c.celestial = sc.texture(config["celestial_texture"], resource_root)
    sc.c.celestial_temperature = float(config["celestial_temperature"])

    pos = camera_position_native(geom, list(opts.camera_position))
    vel = resolve_camera_velocity(geom, pos, config.get("camera_velocity", "StaticObserver"))
    sc.c.camera = make_camera(geom, pos, vel, math.pi / 4.0, opts.height, opts.width, opts.phi, opts.theta, opts.psi)

    for entry in config.get("objects", []):
        (kind, b), = entry.items()
        o = abi.GrrtObject()
        if kind == "Sphere":
            o.kind = abi.OBJ_SPHERE
            o.texture = sc.texture(b["texture"], resource_root)
            o.radius = b["radius"]
            o.position[:] = [float(v) for v in b["position"]]
            o.sphere_temperature = b["temperature"]
        elif kind == "Disc":
            o.kind = abi.OBJ_DISC
            o.texture = sc.texture(b["texture"], resource_root)
            o.inner_radius, o.outer_radius = b["inner_radius"], b["outer_radius"]
            o.temperature = sc.temperature(geom, b["temperature"], b["outer_radius"])
        elif kind == "VolumetricDisc":
            def bad(msg):
                raise OracleError(abi.ERR_INVALID_CONFIGURATION, "VolumetricDisc requires " + msg)
            if b["outer_radius"] <= b["inner_radius"]:
                bad("outer_radius > inner_radius")
            if b["thickness"] <= 0.0:
                bad("thickness > 0")
            if b["max_steps"] == 0:
                bad("max_steps > 0")
            if b["step_size"] <= 0.0:
                bad("step_size > 0")
            if b["brightness_reference_temperature"] <= 0.0:
                bad("brightness_reference_temperature > 0")
            if b["absorption"] < 0.0:
                bad("absorption >= 0")
            if b["scattering"] < 0.0:
                bad("scattering >= 0")
            o.kind = abi.OBJ_VOLUMETRIC_DISC
            o.texture = sc.texture(b["texture"], resource_root)
            o.inner_radius, o.outer_radius = b["inner_radius"], b["outer_radius"]
            o.temperature = sc.temperature(geom, b["temperature"], b["outer_radius"])
            axis = b.get("axis", None) or [0.0, 0.0, 1.0]
            ax, e1, e2 = (C.c_double * 3)(), (C.c_double * 3)(), (C.c_double * 3)()
            lib().orc_volumetric_basis(darr(axis), ax, e1, e2)
            o.axis[:], o.e1[:], o.e2[:] = list(ax), list(e1), list(e2)
            o.num_octaves = b["num_octaves"]
            o.vol_max_steps = b["max_steps"]
            o.vol_step_size = b["step_size"]
            o.thickness = b["thickness"]
            o.density_multiplier = b["density_multiplier"]
            o.brightness_reference_temperature = b["brightness_reference_temperature"]
            o.absorption, o.scattering = b["absorption"], b["scattering"]
            o.noise_scale[:] = [float(v) for v in b["noise_scale"]]
            o.noise_offset = b["noise_offset"]
            perm = (C.c_uint8 * 256)()
            seed = b.get("perlin_seed", None)
            lib().orc_perlin_permutation(1 if seed is None else seed, perm)
            o.perlin_perm[:] = list(perm)
        else:
            raise OracleError(abi.ERR_TOML, f"unknown object {kind}")
        sc.add_object(o)
    return sc


def scene_from_toml(path_or_text, opts=None, resource_root=None, is_text=False):
    opts = opts or abi.default_cli_opts()
    if is_text:
        config = tomllib.loads(path_or_text)
    else:
        with open(path_or_text, "rb") as f:
            config = tomllib.load(f)
    return scene_from_config(config, opts, resource_root)


# ---- running ---------------------------------------------------------------------------

def render(scene, from_row=None, from_col=None, to_row=None, to_col=None, base_only=False):
    """Raytracer::render_section_to_cie_buffer on the CPU oracle."""
    c = scene.c if isinstance(scene, Scene) else scene
    fr = 0 if from_row is None else from_row
    fc = 0 if from_col is None else from_col
    tr = c.camera.rows if to_row is None else to_row
    tc = c.camera.cols if to_col is None else to_col
    n = (tr - fr) * (tc - fc)
    xyza = np.zeros((n, 4))
    cls = np.zeros(n, dtype=np.uint8)
    stop = np.zeros(n, dtype=np.uint8)
    steps = np.zeros(n, dtype=np.uint32)
    stats = abi.GrrtStats()
    rc = lib().orc_render(C.byref(c), fr, fc, tr, tc, dptr(xyza), u8ptr(cls), u8ptr(stop),
                          steps.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(stats), 1 if base_only else 0)
    if rc:
        raise OracleError(rc, "orc_render failed")
    shape = (tr - fr, tc - fc)
    return dict(xyza=xyza.reshape(shape + (4,)), ray_class=cls.reshape(shape), stop=stop.reshape(shape),
                steps=steps.reshape(shape), stats=stats.as_dict())


def color_of_ray(scene, row, col, offset=None):
    c = scene.c if isinstance(scene, Scene) else scene
    out = (C.c_double * 4)()
    cls, stop = C.c_int32(), C.c_int32()
    steps, att = C.c_uint64(), C.c_uint64()
    dx, dy = offset if offset is not None else (0.0, 0.0)
    err = lib().orc_color_of_ray(C.byref(c), row, col, 0 if offset is None else 1, dx, dy, out, C.byref(cls),
                                 C.byref(stop), C.byref(steps), C.byref(att))
    return dict(xyza=list(out), ray_class=cls.value, stop=stop.value, steps=steps.value, attempts=att.value,
                error=err)


def trace_ray(scene, pos=None, mom=None, pixel=None, cap=None):
    """Integrator::integrate with all steps; returns (steps ndarray [n, 10], stop_reason, error)."""
    c = scene.c if isinstance(scene, Scene) else scene
    cap = int(cap or c.integration.max_steps)
    buf = (abi.GrrtStep * cap)()
    n, stop = C.c_uint64(), C.c_int32()
    if pixel is not None:
        err = lib().orc_trace_pixel_ray(C.byref(c), pixel[0], pixel[1], buf, cap, C.byref(n), C.byref(stop))
    else:
        err = lib().orc_trace_ray(C.byref(c), darr(pos), darr(mom), buf, cap, C.byref(n), C.byref(stop))
    m = min(n.value, cap)
    arr = np.frombuffer(buf, dtype=np.dtype([("x", "f8", 4), ("p", "f8", 4), ("t", "f8"), ("step", "u8")]), count=m)
    return arr.copy(), stop.value, err
