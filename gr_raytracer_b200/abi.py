"""ctypes mirror of include/grrt.h (declarations only, no behaviour).

Every Structure below must stay field-for-field identical to the C header; the
`tests/test_abi.py` layout test compares sizes/offsets against what the compiled
library reports.  Reference types mirrored: GeometryType (src/configuration.rs:110-129),
Camera (src/rendering/camera.rs:16-25), IntegrationConfiguration
(src/rendering/integrator.rs:46-67), AdaptiveSamplingConfig (src/configuration.rs:21-58),
GlobalOpts (src/cli/cli.rs:4-48).
"""
import ctypes as C

# enums -------------------------------------------------------------------------------
GEOM_EUCLIDEAN, GEOM_EUCLIDEAN_SPHERICAL, GEOM_SCHWARZSCHILD, GEOM_KERR, GEOM_KERR_BL = range(5)
GEOM_NAMES = ["Euclidean", "EuclideanSpherical", "Schwarzschild", "Kerr", "KerrBL"]
STOP_NONE, STOP_HORIZON, STOP_CELESTIAL, STOP_NAN, STOP_CLOSED_ORBIT, STOP_ERROR = range(6)
STOP_NAMES = ["None", "HorizonReached", "CelestialSphereReached", "CoordinateIsNan",
              "ClosedOrbitDetected", "Error"]
CLASS_ESCAPED, CLASS_CAPTURED, CLASS_HIT = range(3)
TEX_BITMAP, TEX_CHECKER, TEX_BLACKBODY = range(3)
OBJ_SPHERE, OBJ_DISC, OBJ_VOLUMETRIC_DISC = range(3)
TEMP_CONSTANT, TEMP_LUT = range(2)
TONEMAP_REINHARD, TONEMAP_GLOBAL_LINEAR = range(2)
CAMVEL_STATIC, CAMVEL_ZAMO, CAMVEL_EXPLICIT = range(3)
KERR_RHS_REFERENCE_FD = 0

OK = 0
ERR_INVALID_ARGUMENT = -1
ERR_INVALID_CONFIGURATION = -2
ERR_TETRAD_NOT_ORTHONORMAL = -3
ERR_IO = -4
ERR_TOML = -5
ERR_TEXTURE = -6
ERR_CUDA = -7
ERR_OUT_OF_MEMORY = -8
ERR_LUT = -9

MAX_OBJECTS = 8
MAX_PEERS = 8      # GRRT_MAX_PEERS: ranks whose frame buffers one kernel launch can store into
BLACKBODY_LUT_N = 1000
TEMPERATURE_LUT_N = 1000

c_double_p = C.POINTER(C.c_double)
c_uint8_p = C.POINTER(C.c_uint8)


class GrrtGeometry(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("radius", C.c_double),
                ("a", C.c_double), ("horizon_epsilon", C.c_double)]


class GrrtCamera(C.Structure):
    _fields_ = [("position", C.c_double * 4), ("velocity", C.c_double * 4),
                ("tetrad", (C.c_double * 4) * 4), ("alpha", C.c_double),
                ("rows", C.c_int64), ("cols", C.c_int64),
                ("spatial_signature", C.c_double), ("spatial_handedness", C.c_double)]


class GrrtIntegration(C.Structure):
    _fields_ = [("max_steps", C.c_uint64), ("max_radius", C.c_double),
                ("step_size", C.c_double), ("epsilon", C.c_double)]


class GrrtTexture(C.Structure):
    _fields_ = [("kind", C.c_int32), ("width", C.c_uint32), ("height", C.c_uint32),
                ("_pad", C.c_uint32), ("beaming_exponent", C.c_double), ("rgba", c_uint8_p),
                ("checker_width", C.c_double), ("checker_height", C.c_double),
                ("c1", C.c_double * 4), ("c2", C.c_double * 4)]


class GrrtTemperature(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n", C.c_uint32), ("constant", C.c_double),
                ("r_isco", C.c_double), ("lut_r", c_double_p), ("lut_t", c_double_p)]


class GrrtObject(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("texture", GrrtTexture),
                ("temperature", GrrtTemperature),
                ("radius", C.c_double), ("position", C.c_double * 3),
                ("sphere_temperature", C.c_double),
                ("inner_radius", C.c_double), ("outer_radius", C.c_double),
                ("axis", C.c_double * 3), ("e1", C.c_double * 3), ("e2", C.c_double * 3),
                ("num_octaves", C.c_uint64), ("vol_max_steps", C.c_uint64),
                ("vol_step_size", C.c_double), ("thickness", C.c_double),
                ("density_multiplier", C.c_double),
                ("brightness_reference_temperature", C.c_double),
                ("absorption", C.c_double), ("scattering", C.c_double),
                ("noise_scale", C.c_double * 3), ("noise_offset", C.c_double),
                ("perlin_perm", C.c_uint8 * 256)]


class GrrtAdaptive(C.Structure):
    _fields_ = [("enabled", C.c_int32), ("has_minimum_luminance", C.c_int32),
                ("samples_per_axis", C.c_uint64),
                ("luminance_contrast_threshold", C.c_double),
                ("opacity_contrast_threshold", C.c_double),
                ("minimum_luminance", C.c_double),
                ("object_hit_opacity_threshold", C.c_double),
                ("exclude_background_contrast", C.c_int32),
                ("show_sampling_mask", C.c_int32), ("mask_xyza", C.c_double * 4)]


class GrrtScene(C.Structure):
    _fields_ = [("geometry", GrrtGeometry), ("camera", GrrtCamera),
                ("integration", GrrtIntegration), ("adaptive", GrrtAdaptive),
                ("celestial", GrrtTexture), ("celestial_temperature", C.c_double),
                ("n_objects", C.c_uint32), ("kerr_rhs_mode", C.c_int32),
                ("objects", GrrtObject * MAX_OBJECTS),
                ("bb_lut_n", C.c_uint32), ("_pad", C.c_uint32),
                ("bb_lut_log_t", c_double_p), ("bb_lut_xyz", c_double_p)]


class GrrtStep(C.Structure):
    _fields_ = [("x", C.c_double * 4), ("p", C.c_double * 4), ("t", C.c_double),
                ("step", C.c_uint64)]


class GrrtStats(C.Structure):
    _fields_ = [("rays_base", C.c_uint64), ("rays_supersample", C.c_uint64),
                ("pixels_supersampled", C.c_uint64), ("accepted_steps", C.c_uint64),
                ("attempts", C.c_uint64), ("raymarch_samples", C.c_uint64),
                ("stop_histogram", C.c_uint64 * 6), ("ray_errors", C.c_uint64),
                ("hit_stack_overflows", C.c_uint64), ("minimum_luminance", C.c_double),
                ("device_ms_base", C.c_double), ("device_ms_mask", C.c_double),
                ("device_ms_supersample", C.c_double), ("device_ms_total", C.c_double)]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if hasattr(v, "__len__") else v
        return d


class GrrtCliOpts(C.Structure):
    _fields_ = [("width", C.c_int64), ("height", C.c_int64), ("step_size", C.c_double),
                ("max_steps", C.c_uint64), ("max_radius", C.c_double), ("epsilon", C.c_double),
                ("camera_position", C.c_double * 3),
                ("phi", C.c_double), ("theta", C.c_double), ("psi", C.c_double),
                ("tone_mapping", C.c_int32), ("show_sampling_mask", C.c_int32),
                ("sampling_mask_color", C.c_uint8 * 3), ("_pad", C.c_uint8 * 5)]


def default_cli_opts(**overrides):
    """GlobalOpts defaults, src/cli/cli.rs:6-47."""
    o = GrrtCliOpts()
    o.width, o.height = 500, 500
    o.step_size, o.max_steps, o.max_radius, o.epsilon = 0.01, 20000, 15000.0, 0.00001
    o.camera_position[:] = [18.0, 0.0, 0.8]
    o.phi = o.theta = o.psi = 0.0
    o.tone_mapping = TONEMAP_REINHARD
    o.show_sampling_mask = 0
    o.sampling_mask_color[:] = [255, 0, 255]
    for k, v in overrides.items():
        if k in ("camera_position", "sampling_mask_color"):
            getattr(o, k)[:] = list(v)
        else:
            setattr(o, k, v)
    return o
