// grrt_types.h -- plain structs shared by the host-side API code and both kernel translation units.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/grrt.h"

namespace grrt {

constexpr int kTraceThreads = 128;

// ---- device scene (kernel parameter block; lives in the constant bank) ------------------
struct DTexture {
    int kind;
    uint32_t width, height;
    double beaming;
    const uint8_t* rgba;
    double cw, ch;
    double c1[4], c2[4];
};
struct DTemperature {
    int kind;
    uint32_t n;
    double constant, r_isco;
    const double* lut_r;
    const double* lut_t;
};
struct DObject {
    int kind;
    int perm_slot;  // index into the shared-memory Perlin tables
    int axis_is_z;         // VolumetricDisc with the default axis (0,0,1): enables the cheap slab rejection
    double center_norm;    // Sphere: |position|
    DTexture tex;
    DTemperature temp;
    double radius, pos[3], sphere_temperature;
    double r_in, r_out;
    double axis[3], e1[3], e2[3];
    uint64_t num_octaves, vol_max_steps;
    double vol_step, thickness, density_multiplier, bref_temperature, absorption, scattering;
    double noise_scale[3], noise_offset;
    const uint8_t* perm;
};
struct DScene {
    int geom;
    int kerr_mode;
    double radius, a, heps;
    double r_plus_eps;      // horizon radius + epsilon (Kerr / KerrBL), precomputed as the reference does per call
    int horizon_enabled;
    int n_objects;
    // camera
    double cam_pos[4], cam_vel[4], tet[4][4];
    double cam_scale;       // 2 tan(alpha/2) / rows
    double cam_rows, cam_cols, cam_hand, cam_sig;
    // integration
    uint64_t max_steps;
    double max_radius_sq, step_size, epsilon;
    double inv_epsilon;     // 1 / epsilon (fast build: error ratio without a division)
    double hit_opacity_threshold;
    DTexture celestial;
    double celestial_temperature;
    const double* bb_log_t;
    const double* bb_xyz;
    uint32_t bb_n;
    uint32_t n_vol;          // number of VolumetricDisc objects
    const double* srgb_lut;  // 256 entries: inv_compand_srgb(i/255) (color.rs:301-308), built on the host
    // summary of the objects for segment_needs_cartesian() (spherical / Boyer-Lindquist charts), built at upload:
    // straight-line tests instead of a loop over the objects with a switch on their kind
    int pre_flags;           // PRE_* bits
    int pre_n_spheres;
    double pre_vol_h;        // largest capture half-height 3 thickness (1 + 1e-9) of the VolumetricDiscs
    double pre_sph_far2[GRRT_MAX_OBJECTS], pre_sph_near2[GRRT_MAX_OBJECTS];   // near2 < 0: the sphere contains the origin
    DObject obj[GRRT_MAX_OBJECTS];
};
enum { PRE_HAS_DISC = 1, PRE_HAS_VOL = 2, PRE_ALWAYS = 4 /* a VolumetricDisc with a tilted axis: no cheap rejection */ };

enum { CNT_ATTEMPTS = 0, CNT_ACCEPTED = 1, CNT_SAMPLES = 2, CNT_ERRORS = 3, CNT_OVERFLOW = 4, CNT_RAYS = 5,
       CNT_STOP0 = 8 /* 8..13 */ };

struct TraceArgs {
    const uint32_t* items;      // nullable: pixel ids (row-major within the rectangle); item i is items[item_offset + i * item_stride]
    unsigned long long item_offset, item_stride;   // a rank's share of a list that is dealt round-robin (stride >= 1)
    unsigned long long first_item;
    unsigned long long n_items;
    unsigned long long n_units;  // work units handed out by the queue
    uint32_t from_row, from_col, width, height;
    uint32_t spa;                // 0: centre ray per item;  N: N*N stratified sub-samples per item
    uint32_t tiled;              // items == NULL && spa == 0: units enumerate 8x4 pixel tiles
    uint32_t tiles_x;
    double* xyza;                // spa == 0: per item;  spa > 0: per unit
    uint8_t* cls;
    uint8_t* stop;
    uint32_t* steps;
    uint8_t* sub_valid;          // spa > 0: per unit
    unsigned long long* queue;
    unsigned long long* counters;
    // n_peers > 0 (spa == 0 only): results are written at the item's PIXEL id into every peer's frame buffers
    // (peer-mapped device pointers, this GPU's own included) instead of compactly into xyza / cls / stop
    uint32_t n_peers;
    double* xyza_p[GRRT_MAX_PEERS];
    uint8_t* cls_p[GRRT_MAX_PEERS];
    uint8_t* stop_p[GRRT_MAX_PEERS];
    uint32_t* steps_p[GRRT_MAX_PEERS];   // nullable entries
};


// implemented in grrt_kern_exact.cu (-fmad=false) and grrt_kern_fast.cu (-fmad=true)
struct TraceLaunch {
    int num_sms;
    cudaStream_t stream;
};
}  // namespace grrt

namespace grrt_exact {
int launch_trace(const grrt::DScene& s, const grrt::TraceArgs& a, bool has_vol, const grrt::TraceLaunch& l);
int launch_trace_single(const grrt::DScene& s, const double pos[4], const double mom[4], int use_pixel, long long row, long long col,
                        GrrtStep* d_steps, unsigned long long cap, unsigned long long* d_n, int* d_stop, int* d_err,
                        cudaStream_t stream);
int launch_debug_pow(const double* x, const double* y, double* out, size_t n, cudaStream_t stream);
int launch_debug_sincos(const double* x, double* sn, double* cs, size_t n, cudaStream_t stream);
}
namespace grrt_fast {
int launch_trace(const grrt::DScene& s, const grrt::TraceArgs& a, bool has_vol, const grrt::TraceLaunch& l);
}
