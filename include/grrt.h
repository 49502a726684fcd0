/*
 * grrt.h -- C ABI of libgrrt, the B200 (sm_100a) implementation of gr_raytracer's
 * per-pixel geodesic-integration path.
 *
 * The reference (mdreem/gr_raytracer, pure Rust, CPU/rayon) has no FFI seam; the cut
 * chosen here is the one SURVEY.md 8(b) derives from its call graph:
 *
 *   Raytracer::render_section_to_cie_buffer   src/rendering/raytracer.rs:177-193
 *   Raytracer::integrate_ray_at_point         src/rendering/raytracer.rs:499-507
 *   integrate_and_save_ray                    src/cli/shared.rs:107-129
 *   xyz_to_linear_srgb_buffer / linear_srgb_to_srgb_buffer   src/rendering/color.rs:236-299
 *
 * Everything the device needs is immutable after `create_scene`
 * (src/cli/shared.rs:131-321); it is handed over as the plain-old-data GrrtScene below.
 * All structs are #[repr(C)]-compatible: fixed-width integers, doubles, raw pointers,
 * no torch / C++ types.  Ownership: the caller owns every pointer it passes in (the
 * scene is deep-copied to the device inside the call); the library owns GrrtCtx and all
 * device memory.  Errors: integer return codes (0 = ok, <0 = GrrtStatus), no exceptions
 * cross the boundary; per-ray failures are not API errors, they reproduce the
 * reference's "log and keep the default pixel" (raytracer.rs:230-239) and are counted
 * in GrrtStats.  A GrrtCtx is NOT re-entrant: one in-flight call per context.  That includes
 * the *_device entry points that take a caller stream -- they share per-context scratch (the
 * work-queue counter, the sort / select temporary storage, the tone-map maximum), so all
 * calls on one context must be issued from one thread at a time and onto one stream (or be
 * ordered by the caller with events).  Use one GrrtCtx per concurrent stream.
 */
#ifndef GRRT_H
#define GRRT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GRRT_MAX_PEERS 8   /* devices of one multi-device context / frames one kernel launch can store into */

/* ---- enums (values are ABI) ------------------------------------------------------ */

/* configuration.rs:110-129 GeometryType */
enum {
    GRRT_GEOM_EUCLIDEAN = 0,
    GRRT_GEOM_EUCLIDEAN_SPHERICAL = 1,
    GRRT_GEOM_SCHWARZSCHILD = 2,
    GRRT_GEOM_KERR = 3,      /* Kerr-Schild Cartesian chart, geometry/kerr.rs */
    GRRT_GEOM_KERR_BL = 4    /* Boyer-Lindquist, Carter-separated, geometry/kerr_bl.rs */
};

/* integrator.rs:21-27 StopReason; NONE = step budget ran out without a terminal event */
enum {
    GRRT_STOP_NONE = 0,
    GRRT_STOP_HORIZON = 1,
    GRRT_STOP_CELESTIAL = 2,
    GRRT_STOP_NAN = 3,
    GRRT_STOP_CLOSED_ORBIT = 4,
    GRRT_STOP_ERROR = 5      /* the ray returned Err(..): pixel keeps its default value */
};

/* scene.rs:25-30 RayClass */
enum { GRRT_CLASS_ESCAPED = 0, GRRT_CLASS_CAPTURED = 1, GRRT_CLASS_HIT = 2 };

/* per-ray error causes (the RaytracerError variants that can arise inside color_of_ray) */
enum {
    GRRT_RAYERR_NONE = 0,
    GRRT_RAYERR_RKF45_RETRIES = 1,      /* runge_kutta.rs:179-181 */
    GRRT_RAYERR_NO_CIRCULAR_ORBIT = 2,  /* circular_orbit.rs:93-101 via disc.rs:107 */
    GRRT_RAYERR_BELOW_RISCO = 3,        /* temperature.rs:204-217 */
    GRRT_RAYERR_NONFINITE_RADIUS = 4    /* temperature.rs:199-202 */
};

enum { GRRT_TEX_BITMAP = 0, GRRT_TEX_CHECKER = 1, GRRT_TEX_BLACKBODY = 2 };
enum { GRRT_OBJ_SPHERE = 0, GRRT_OBJ_DISC = 1, GRRT_OBJ_VOLUMETRIC_DISC = 2 };
enum { GRRT_TEMP_CONSTANT = 0, GRRT_TEMP_LUT = 1 };
enum { GRRT_TONEMAP_REINHARD = 0, GRRT_TONEMAP_GLOBAL_LINEAR = 1 };
enum { GRRT_CAMVEL_STATIC = 0, GRRT_CAMVEL_ZAMO = 1, GRRT_CAMVEL_EXPLICIT = 2 };

/* GrrtScene.kerr_rhs_mode: the Kerr-Schild right-hand side differentiates the metric by central differences exactly as
 * kerr.rs:162-186 does (the only mode; grrt_render / grrt_scene_upload reject any other value). */
enum { GRRT_KERR_RHS_REFERENCE_FD = 0 };

typedef enum {
    GRRT_OK = 0,
    GRRT_ERR_INVALID_ARGUMENT = -1,
    GRRT_ERR_INVALID_CONFIGURATION = -2,  /* RaytracerError::InvalidConfiguration */
    GRRT_ERR_TETRAD_NOT_ORTHONORMAL = -3, /* CameraError::TetradNotOrthonormal */
    GRRT_ERR_IO = -4,                     /* IoError / ConfigurationFileError */
    GRRT_ERR_TOML = -5,                   /* TomlError */
    GRRT_ERR_TEXTURE = -6,                /* TextureError */
    GRRT_ERR_CUDA = -7,                   /* no usable device / launch failure: never a CPU fallback */
    GRRT_ERR_OUT_OF_MEMORY = -8,
    GRRT_ERR_LUT = -9                     /* DenominatorCloseToZero / NoCircularOrbit while building a LUT */
} GrrtStatus;

/* ---- scene description ------------------------------------------------------------ */

#define GRRT_MAX_OBJECTS 8
#define GRRT_BLACKBODY_LUT_N 1000   /* texture.rs:116 NUM_COLOR_LUT_STEPS */
#define GRRT_TEMPERATURE_LUT_N 1000 /* temperature.rs:40 NUM_LUT_STEPS */

typedef struct {
    int32_t kind;      /* GRRT_GEOM_* */
    int32_t _pad;
    double radius;     /* r_s = 2M */
    double a;          /* spin */
    double horizon_epsilon;
} GrrtGeometry;

/* rendering/camera.rs:16-25.  position/velocity are in the geometry's native chart
 * (spherical for Schwarzschild/EuclideanSpherical, Kerr-Schild Cartesian for Kerr,
 * Boyer-Lindquist for KerrBL, Cartesian for Euclidean).  tetrad rows are e_t, e_x, e_y,
 * e_z AFTER the phi/theta/psi rotations and the boost to `velocity`. */
typedef struct {
    double position[4];
    double velocity[4];
    double tetrad[4][4];
    double alpha;              /* cli/shared.rs:162: PI/4 */
    int64_t rows, cols;        /* rows = --height, cols = --width */
    double spatial_signature;  /* signature[3] */
    double spatial_handedness; /* camera.rs:134-148 */
} GrrtCamera;

/* rendering/integrator.rs:46-67 */
typedef struct {
    uint64_t max_steps;
    double max_radius;
    double step_size;
    double epsilon;
} GrrtIntegration;

/* rendering/texture.rs: TextureMapper (bitmap) / CheckerMapper / BlackBodyMapper */
typedef struct {
    int32_t kind;              /* GRRT_TEX_* */
    uint32_t width, height;    /* bitmap size in texels */
    uint32_t _pad;
    double beaming_exponent;
    const uint8_t* rgba;       /* bitmap: width*height*4 bytes, row-major, 8-bit sRGB + alpha */
    double checker_width, checker_height;
    double c1[4], c2[4];       /* checker: CIE XYZ + alpha of color1 / color2 (srgb_to_xyz) */
} GrrtTexture;

/* rendering/temperature.rs: ConstantTemperatureComputer | KerrTemperatureComputer */
typedef struct {
    int32_t kind;              /* GRRT_TEMP_* */
    uint32_t n;                /* LUT entries */
    double constant;           /* GRRT_TEMP_CONSTANT */
    double r_isco;             /* GRRT_TEMP_LUT: radii below this are BelowRISCO errors */
    const double* lut_r;       /* n radii, ascending */
    const double* lut_t;       /* n temperatures */
} GrrtTemperature;

/* scene_objects/{sphere,disc,volumetric_disc}.rs */
typedef struct {
    int32_t kind;              /* GRRT_OBJ_* */
    int32_t _pad;
    GrrtTexture texture;
    GrrtTemperature temperature; /* Disc / VolumetricDisc */
    /* Sphere */
    double radius;
    double position[3];
    double sphere_temperature;
    /* Disc / VolumetricDisc */
    double inner_radius, outer_radius;
    /* VolumetricDisc (volumetric_disc.rs:21-40) */
    double axis[3], e1[3], e2[3];  /* normalised axis and in-plane basis (ctor :61-73) */
    uint64_t num_octaves;
    uint64_t vol_max_steps;
    double vol_step_size;
    double thickness;
    double density_multiplier;
    double brightness_reference_temperature;
    double absorption, scattering;
    double noise_scale[3];
    double noise_offset;
    uint8_t perlin_perm[256];      /* noise 0.9.0 PermutationTable::new(perlin_seed) */
} GrrtObject;

/* configuration.rs:21-58 AdaptiveSamplingConfig + the mask colour of cli.rs:41-47 */
typedef struct {
    int32_t enabled;
    int32_t has_minimum_luminance;
    uint64_t samples_per_axis;
    double luminance_contrast_threshold;
    double opacity_contrast_threshold;
    double minimum_luminance;
    double object_hit_opacity_threshold;
    int32_t exclude_background_contrast;
    int32_t show_sampling_mask;
    double mask_xyza[4];
} GrrtAdaptive;

typedef struct {
    GrrtGeometry geometry;
    GrrtCamera camera;
    GrrtIntegration integration;
    GrrtAdaptive adaptive;
    GrrtTexture celestial;
    double celestial_temperature;
    uint32_t n_objects;
    int32_t kerr_rhs_mode;         /* must be GRRT_KERR_RHS_REFERENCE_FD (0) */
    GrrtObject objects[GRRT_MAX_OBJECTS];
    /* BlackBodyMapper LUT (texture.rs:121-138), shared by every BlackBody texture */
    uint32_t bb_lut_n;
    uint32_t _pad;
    const double* bb_lut_log_t;    /* bb_lut_n keys: log10(T) */
    const double* bb_lut_xyz;      /* bb_lut_n * 3 */
} GrrtScene;

/* One stored integrator step (integrator.rs:13-19) for trajectory export. */
typedef struct {
    double x[4];   /* position, native chart */
    double p[4];   /* contravariant momentum, native chart */
    double t;      /* integration parameter (affine, or Mino time for KerrBL) */
    uint64_t step;
} GrrtStep;

typedef struct {
    uint64_t rays_base;            /* 1-spp rays traced */
    uint64_t rays_supersample;     /* sub-sample rays traced */
    uint64_t pixels_supersampled;
    uint64_t accepted_steps;       /* sum over rays of (stored steps - 1) */
    uint64_t attempts;             /* rkf45_step evaluations */
    uint64_t raymarch_samples;     /* VolumetricDisc samples */
    uint64_t stop_histogram[6];    /* indexed by GRRT_STOP_* */
    uint64_t ray_errors;
    uint64_t hit_stack_overflows;
    double minimum_luminance;      /* resolved faintness floor (raytracer.rs:118-129) */
    double device_ms_base;         /* CUDA-event time of the base pass */
    double device_ms_mask;
    double device_ms_supersample;
    double device_ms_total;
} GrrtStats;

typedef struct GrrtCtx GrrtCtx;
typedef struct GrrtDeviceScene GrrtDeviceScene;

/* ---- device entry points (libgrrt.so, CUDA) ------------------------------------- */

/* Create a context bound to one CUDA device.  Fails with GRRT_ERR_CUDA when there is
 * no usable sm_100 device: there is no CPU fallback. */
int grrt_init(int device_id, GrrtCtx** out);
/* One context over n_devices (1..GRRT_MAX_PEERS, distinct) GPUs of one node driven by this process -- the
 * grrt_init(n_devices, device_ids) of SURVEY.md 8(b).  The reference's entry point is one blocking call of one process
 * (src/main.rs:106-115 -> raytracer.rs:460-497); grrt_render / grrt_render_resident on such a context are likewise
 * blocking and shard the frame internally: 8x4-pixel tiles are dealt round-robin to the devices, every device stores
 * its finished pixels straight into the frame held by device_ids[0] (peer stores over NVLink from inside the trace
 * kernel; there is no reduction, so no collective and no gather pass), the flagged pixels are dealt in descending
 * base-pass cost.  The frame is bit for bit the single-device frame.  Needs peer access from every device to
 * device_ids[0] (GRRT_ERR_CUDA otherwise).  Scenes must be uploaded through the same context. */
int grrt_init_multi(int n_devices, const int* device_ids, GrrtCtx** out);
int grrt_device_count(const GrrtCtx* ctx);   /* 1 for grrt_init contexts */
void grrt_destroy(GrrtCtx* ctx);
const char* grrt_last_error(const GrrtCtx* ctx);
const char* grrt_version(void);
/* Options: "fma" (default 1): the smooth geometries (Euclidean, EuclideanSpherical, Schwarzschild, KerrBL) run kernels
 * compiled with FMA contraction; 0 selects the contraction-free kernels for every geometry (Kerr-Schild always
 * uses those).  Environment variable GRRT_FMA=0 has the same effect at grrt_init. */
int grrt_set_option(GrrtCtx* ctx, const char* name, int value);
/* Number of this library's own kernels launched through `ctx` so far (CUB/memset launches not counted). */
unsigned long long grrt_kernel_launch_count(const GrrtCtx* ctx);

/* Drop-in for Raytracer::render_section_to_cie_buffer (raytracer.rs:177-193): base
 * pass, adaptive mask, supersampling (or mask visualisation).  Outputs are caller-owned
 * HOST buffers of (to_row-from_row)*(to_col-from_col) pixels, row-major; all but
 * xyza_out may be NULL. */
int grrt_render(GrrtCtx* ctx, const GrrtScene* scene,
                uint32_t from_row, uint32_t from_col, uint32_t to_row, uint32_t to_col,
                double* xyza_out, uint8_t* ray_class_out, uint8_t* stop_reason_out,
                uint32_t* steps_out, GrrtStats* stats);

/* xyz_to_linear_srgb_buffer + linear_srgb_to_srgb_buffer (color.rs:236-299). */
int grrt_tonemap(GrrtCtx* ctx, const double* xyza, size_t n_pixels, int method,
                 double exposure, uint8_t* rgb_out);

/* Integrator::integrate for one ray with every stored step written out
 * (raytracer.rs:499-507, cli/shared.rs:107-129).  pos/mom: native chart, contravariant. */
int grrt_trace_ray(GrrtCtx* ctx, const GrrtScene* scene, const double pos[4],
                   const double mom[4], GrrtStep* steps_out, uint64_t cap,
                   uint64_t* n_out, int32_t* stop_reason);
/* Same for the camera ray of pixel (row, col) (Camera::get_ray_for, camera.rs:234-245). */
int grrt_trace_pixel_ray(GrrtCtx* ctx, const GrrtScene* scene, int64_t row, int64_t col,
                         GrrtStep* steps_out, uint64_t cap, uint64_t* n_out,
                         int32_t* stop_reason);

/* Device-resident variants: the scene is uploaded once, inputs/outputs are DEVICE
 * pointers and the work is enqueued on `cuda_stream` (a cudaStream_t passed as void*)
 * without host synchronisation.  These are what bench.py's `value` and the multi-GPU
 * tile sharding use. */
int grrt_scene_upload(GrrtCtx* ctx, const GrrtScene* scene, GrrtDeviceScene** out);
void grrt_scene_free(GrrtCtx* ctx, GrrtDeviceScene* dscene);

/* Animation frames (the reference re-runs its binary per camera position,
 * scripts/Readme.md "Camera fly-through movie"): keep textures and look-up tables
 * resident, replace only the camera (Camera::new output, camera.rs:16-25), and render
 * the resident scene into HOST buffers exactly as grrt_render does.  The frame equals
 * grrt_render of a scene assembled with that camera, bit for bit. */
int grrt_scene_set_camera(GrrtCtx* ctx, GrrtDeviceScene* dscene, const GrrtCamera* camera);
int grrt_render_resident(GrrtCtx* ctx, const GrrtDeviceScene* dscene,
                         uint32_t from_row, uint32_t from_col, uint32_t to_row, uint32_t to_col,
                         double* xyza_out, uint8_t* ray_class_out, uint8_t* stop_reason_out,
                         uint32_t* steps_out, GrrtStats* stats);

/* Trace `n_items` work items.  Item i is pixel id items[i] of the rectangle
 * [from_row,to_row) x [from_col,to_col), numbered row-major from its top-left corner
 * (items == NULL: the consecutive ids first_item .. first_item+n_items-1).  samples_per_axis == 0:
 * one centre ray per item.  samples_per_axis == N > 0: N*N stratified sub-sample rays
 * per item (raytracer.rs:138-159) averaged over the valid ones; items whose sub-samples
 * all failed keep xyza untouched.  Outputs are indexed by i. */
int grrt_trace_items_device(GrrtCtx* ctx, const GrrtDeviceScene* dscene,
                            const uint32_t* d_items, uint64_t first_item, uint64_t n_items,
                            uint32_t from_row, uint32_t from_col, uint32_t to_row, uint32_t to_col,
                            uint32_t samples_per_axis,
                            double* d_xyza, uint8_t* d_class, uint8_t* d_stop,
                            uint32_t* d_steps, uint64_t* d_counters /* 16 u64, accumulated */,
                            void* cuda_stream);

/* Multi-GPU frames without a gather pass (tiles travel by peer stores, there
 * is no reduction).  Every rank allocates its frame buffers with grrt_ipc_alloc, the 64-byte
 * handles are exchanged by the host (e.g. torch.distributed), grrt_ipc_open maps the peers'
 * buffers, and grrt_trace_items_device_peers writes each finished pixel at its PIXEL id into
 * all `n_peers` frames (own pointer included) straight from the trace kernel / the sub-sample
 * mean kernel.  The pointer arrays are HOST arrays of device pointers.  The caller orders the
 * passes across ranks (a barrier after each call) before reading its frame. */
int grrt_ipc_alloc(GrrtCtx* ctx, size_t bytes, void** d_ptr, unsigned char handle_out[64]);
int grrt_ipc_open(GrrtCtx* ctx, const unsigned char handle[64], void** d_ptr);
int grrt_ipc_close(GrrtCtx* ctx, void* d_ptr);
int grrt_ipc_free(GrrtCtx* ctx, void* d_ptr);
/* Item i of the call is d_items[item_offset + i * item_stride] (item_stride >= 1): a rank's share of a list
 * that is dealt round-robin.  d_steps_peers (and each of its entries) may be NULL. */
int grrt_trace_items_device_peers(GrrtCtx* ctx, const GrrtDeviceScene* dscene,
                                  const uint32_t* d_items, uint64_t item_offset, uint64_t item_stride,
                                  uint64_t n_items,
                                  uint32_t from_row, uint32_t from_col, uint32_t to_row, uint32_t to_col,
                                  uint32_t samples_per_axis, uint32_t n_peers,
                                  double* const* d_xyza_peers, uint8_t* const* d_class_peers,
                                  uint8_t* const* d_stop_peers, uint32_t* const* d_steps_peers,
                                  uint64_t* d_counters, void* cuda_stream);

/* The base pass as pilot of the supersampling pass (BASELINE.json north_star: "load-balanced from a pilot ...
 * pass"): orders the flagged pixel ids by descending base-pass step count (stable, ties keep pixel order).
 * Tracing the list in that order ends the work queue on short rays; dealing it round-robin gives every
 * device the same share of the expensive pixels.  The image does not depend on the order. */
int grrt_order_by_cost_device(GrrtCtx* ctx, const uint32_t* d_selected, uint32_t n_selected,
                              const uint32_t* d_steps /* per pixel id */, uint32_t* d_ordered, void* cuda_stream);

/* resolve_minimum_luminance + collect_pixels_to_supersample (raytracer.rs:118-129,
 * 386-458) over a complete base buffer of rows x cols pixels on the device.  Writes the
 * flagged pixel ids (ascending, i.e. the reference's row-major push order) to
 * d_selected and their count to *d_count. */
int grrt_select_mask_device(GrrtCtx* ctx, const GrrtDeviceScene* dscene,
                            const double* d_xyza, const uint8_t* d_class,
                            uint32_t rows, uint32_t cols,
                            uint32_t* d_selected, uint32_t* d_count,
                            double* minimum_luminance_out, void* cuda_stream);

int grrt_tonemap_device(GrrtCtx* ctx, const double* d_xyza, size_t n_pixels, int method,
                        double exposure, uint8_t* d_rgb, void* cuda_stream);

/* DFMA-chain microbenchmark: the measured FP64 issue-rate roofline denominator.
 * Returns FP64 warp-level lane-instructions per second (one DFMA = 1). */
int grrt_measure_fp64_peak(GrrtCtx* ctx, double* dfma_per_second, double* ms);

/* Test hook: evaluates the device's glibc-compatible pow (csrc/grrt_pow.cuh) on host arrays. */
int grrt_debug_pow(GrrtCtx* ctx, const double* x, const double* y, double* out, size_t n);
/* Test hook: the kernels' sin/cos routine on host arrays. */
int grrt_debug_sincos(GrrtCtx* ctx, const double* x, double* sin_out, double* cos_out, size_t n);

/* ---- host-side scene assembly (mirror of src/cli + src/configuration.rs) --------- */

/* cli.rs:4-48 GlobalOpts */
typedef struct {
    int64_t width, height;
    double step_size;
    uint64_t max_steps;
    double max_radius;
    double epsilon;
    double camera_position[3];
    double phi, theta, psi;
    int32_t tone_mapping;        /* GRRT_TONEMAP_* */
    int32_t show_sampling_mask;
    uint8_t sampling_mask_color[3];
    uint8_t _pad[5];
} GrrtCliOpts;

typedef struct GrrtHostScene GrrtHostScene;

void grrt_cli_opts_default(GrrtCliOpts* opts);
/* TOML file -> RenderConfig -> create_scene (cli/shared.rs:131-321) -> GrrtScene.
 * `resource_root` (nullable) is prepended to relative texture paths. */
int grrt_host_scene_from_toml_file(const char* toml_path, const char* resource_root,
                                   const GrrtCliOpts* opts, GrrtHostScene** out,
                                   char* err, size_t err_len);
int grrt_host_scene_from_toml_string(const char* toml_text, const char* resource_root,
                                     const GrrtCliOpts* opts, GrrtHostScene** out,
                                     char* err, size_t err_len);
const GrrtScene* grrt_host_scene_get(const GrrtHostScene* hs);
void grrt_host_scene_free(GrrtHostScene* hs);
/* Recomputes only the camera of an assembled scene (position, phi/theta/psi, width and
 * height of `opts`; cli/shared.rs:23-86 + Camera::new): one call per animation frame. */
int grrt_host_scene_set_camera(GrrtHostScene* hs, const GrrtCliOpts* opts, char* err, size_t err_len);
/* The `blackbody` / `blackbody-spectrum` utilities (cli/blackbody.rs): CIE XYZ of a red-shifted
 * black body (black_body_radiation.rs:12-41, the integral the black-body LUT is built from) and
 * xyz_to_srgb (color.rs:218-233: matrix, exposure, clamp, companding; no tone mapping). */
void grrt_host_blackbody_xyz(double temperature, double redshift, double xyz_out[3]);
void grrt_host_xyz_to_srgb8(const double xyz[3], double exposure, uint8_t rgb_out[3]);
/* render-ray-at: momentum from a Cartesian direction in the unboosted tetrad
 * (cli/schwarzschild.rs:89-121, cli/kerr.rs:77-100, cli/kerr_bl.rs:78-107). */
int grrt_host_ray_at(const GrrtGeometry* geometry, const double position_cartesian[3],
                     const double direction_cartesian[3], double pos_out[4], double mom_out[4],
                     char* err, size_t err_len);
/* IntegratedRay::save (ray.rs:35-54): "i,t,tau,x,y,z" CSV. */
int grrt_host_write_ray_csv(const GrrtGeometry* geometry, const GrrtStep* steps, uint64_t n,
                            const char* path);
int grrt_host_write_png_rgb8(const char* path, const uint8_t* rgb, uint32_t width, uint32_t height);
int grrt_host_write_hdr_xyz(const char* path, const double* xyza, uint32_t width, uint32_t height);
int grrt_host_read_png_rgba8(const char* path, uint8_t** rgba_out, uint32_t* width, uint32_t* height);
void grrt_host_free(void* p);
/* sizeof() of the ABI structs in declaration order (0 = GrrtGeometry ... 10 = GrrtCliOpts), for binding checks */
size_t grrt_abi_sizeof(int which);

#ifdef __cplusplus
}
#endif
#endif /* GRRT_H */
