// grrt_api.cu -- device half of the C ABI declared in include/grrt.h.
//
// Compiled for sm_100a only (-gencode arch=compute_100a,code=sm_100a -fmad=false).  There is no CPU
// path in this library: every entry point that computes needs a CUDA device and fails with
// GRRT_ERR_CUDA otherwise.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <cub/cub.cuh>

#include "grrt_types.h"

using namespace grrt;

struct GrrtCtx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;    // device -> host copy of the base frame while the supersampling pass runs
    void* h_stage = nullptr;               // pinned staging for the supersampled pixels (ids + values)
    size_t h_stage_bytes = 0;
    std::string last_error;
    double* d_srgb_lut = nullptr;
    unsigned long long* d_queue = nullptr;     // work-queue counter
    unsigned long long* d_counters = nullptr;  // 16 x u64
    void* d_cub_temp = nullptr;
    size_t cub_temp_bytes = 0;
    unsigned long long launches = 0;  // kernels of this library launched so far (bench.py's gpu_launches)
    bool use_fma = true;              // smooth geometries: FMA-contracted kernels (GRRT_FMA=0 or grrt_set_option to disable)
    // grrt_init_multi: the context of device_ids[0] owns one member context per further device; frames are sharded
    // over all of them and assembled in this (the primary) device's memory by peer stores
    std::vector<GrrtCtx*> members;
    // per-device share of the 8x4-tile round-robin partition of the last frame geometry rendered (multi-device only)
    uint32_t part_w = 0, part_h = 0;
    std::vector<uint32_t*> d_part_items;   // [1 + members.size()], each on its own device
    std::vector<uint64_t> part_counts;
};

struct GrrtDeviceScene {
    DScene s;
    std::vector<void*> allocs;
    std::vector<size_t> alloc_bytes;   // parallel to allocs
    GrrtAdaptive adaptive;
    bool has_vol = false;
    std::vector<GrrtDeviceScene*> replicas;  // multi-device context: the same scene on members[k]'s device
};

namespace {

#define CUDA_TRY(ctx, expr)                                                                     \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            (ctx)->last_error = std::string(#expr) + ": " + cudaGetErrorString(_e);             \
            return _e == cudaErrorMemoryAllocation ? GRRT_ERR_OUT_OF_MEMORY : GRRT_ERR_CUDA;    \
        }                                                                                       \
    } while (0)

int fail(GrrtCtx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->last_error = msg;
    return code;
}

double inv_compand_srgb_host(double u) {  // color.rs:301-308 (glibc pow, as the reference)
    if (u <= 0.04045) return u / 12.92;
    return std::pow((u + 0.055) / 1.055, 2.4);
}

template <class T>
int upload(GrrtCtx* ctx, GrrtDeviceScene* ds, const T* host, size_t count, const T** out) {
    *out = nullptr;
    if (!host || count == 0) return GRRT_OK;
    void* p = nullptr;
    CUDA_TRY(ctx, cudaMallocAsync(&p, count * sizeof(T), ctx->stream));
    ds->allocs.push_back(p);
    ds->alloc_bytes.push_back(count * sizeof(T));
    CUDA_TRY(ctx, cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    *out = static_cast<const T*>(p);
    return GRRT_OK;
}

int convert_texture(GrrtCtx* ctx, GrrtDeviceScene* ds, const GrrtTexture& t, DTexture* o,
                    std::map<const uint8_t*, const uint8_t*>& cache) {
    std::memset(o, 0, sizeof *o);
    o->kind = t.kind;
    o->width = t.width;
    o->height = t.height;
    o->beaming = t.beaming_exponent;
    o->cw = t.checker_width;
    o->ch = t.checker_height;
    for (int i = 0; i < 4; ++i) {
        o->c1[i] = t.c1[i];
        o->c2[i] = t.c2[i];
    }
    if (t.kind == GRRT_TEX_BITMAP) {
        if (!t.rgba || t.width == 0 || t.height == 0) return fail(ctx, GRRT_ERR_INVALID_ARGUMENT, "bitmap texture without pixels");
        auto it = cache.find(t.rgba);
        if (it == cache.end()) {
            const uint8_t* d = nullptr;
            int rc = upload<uint8_t>(ctx, ds, t.rgba, (size_t)t.width * t.height * 4, &d);
            if (rc) return rc;
            it = cache.emplace(t.rgba, d).first;
        }
        o->rgba = it->second;
    }
    return GRRT_OK;
}

// ---- adaptive-sampling mask (raytracer.rs:59-129, 386-458) ------------------------------------------
__global__ void luminance_keys_kernel(const double* xyza, size_t n, unsigned long long* keys) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long b = __double_as_longlong(xyza[4 * i + 1]);
    // f64::total_cmp order as an unsigned key
    unsigned long long u = (unsigned long long)b;
    u = (b < 0) ? ~u : (u | 0x8000000000000000ull);
    keys[i] = u;
}
__device__ __forceinline__ bool should_supersample_pair(const double* p, int pc, const double* q, int qc,
                                                        const GrrtAdaptive& cfg, double minimum_luminance) {
    if (pc != qc) return true;
    if (cfg.exclude_background_contrast && pc == GRRT_CLASS_ESCAPED) return false;
    bool visible = fmax(p[1], q[1]) > minimum_luminance;
    double lc = fabs(p[1] - q[1]) / (p[1] + q[1] + 1e-4);
    double oc = fabs(p[3] - q[3]);
    return visible && (lc > cfg.luminance_contrast_threshold || oc > cfg.opacity_contrast_threshold);
}
__global__ void mask_flags_kernel(const double* xyza, const uint8_t* cls, uint32_t H, uint32_t W,
                                  const __grid_constant__ GrrtAdaptive cfg, const unsigned long long* key_at_index,
                                  double* min_lum_out, uint8_t* flags) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    double min_lum;
    if (cfg.has_minimum_luminance) {
        min_lum = cfg.minimum_luminance;
    } else {
        unsigned long long u = *key_at_index;
        long long b = (u & 0x8000000000000000ull) ? (long long)(u & 0x7fffffffffffffffull) : (long long)~u;
        min_lum = 1e-3 * __longlong_as_double(b);  // RELATIVE_MINIMUM_LUMINANCE_FRACTION
    }
    if (i == 0) *min_lum_out = min_lum;
    if (i >= (size_t)H * W) return;
    int row = (int)(i / W), col = (int)(i % W);
    const double* p = xyza + 4 * i;
    int pc = cls[i];
    bool flag = false;
    const int SH[8][2] = {{-1, -1}, {-1, 0}, {-1, 1}, {0, -1}, {0, 1}, {1, -1}, {1, 0}, {1, 1}};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        int nr = row + SH[k][0], nc = col + SH[k][1];
        if (nr < 0 || nr >= (int)H || nc < 0 || nc >= (int)W) continue;
        size_t ni = (size_t)nr * W + nc;
        if (should_supersample_pair(p, pc, xyza + 4 * ni, cls[ni], cfg, min_lum)) {
            flag = true;
            break;
        }
    }
    flags[i] = flag ? 1 : 0;
}
__global__ void fill_mask_color_kernel(const uint32_t* sel, uint32_t n, const __grid_constant__ GrrtAdaptive cfg, double* xyza) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double* d = xyza + 4ull * sel[i];
    d[0] = cfg.mask_xyza[0]; d[1] = cfg.mask_xyza[1]; d[2] = cfg.mask_xyza[2]; d[3] = cfg.mask_xyza[3];
}

__device__ __forceinline__ double rclamp(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

// ---- tone mapping (color.rs:193-299) ----------------------------------------------------------------
__device__ __forceinline__ void xyz_to_linear(const double* c, double out[3]) {
    out[0] = (3.2406255 * c[0] + -1.5372080 * c[1]) + -0.4986286 * c[2];
    out[1] = (-0.9689307 * c[0] + 1.8757561 * c[1]) + 0.0415175 * c[2];
    out[2] = (0.0557101 * c[0] + -0.2040211 * c[1]) + 1.0569959 * c[2];
}
__device__ __forceinline__ double compand_srgb(double linear) {
    double sign = linear < 0.0 ? -1.0 : 1.0;
    double a = fabs(linear);
    double enc = (a <= 0.0031308) ? 12.92 * a : 1.055 * pow(a, 1.0 / 2.4) - 0.055;
    return rclamp(sign * enc, 0.0, 1.0);
}
__device__ __forceinline__ uint8_t to_u8(double v) {
    double r = round(v);
    if (!(r > 0.0)) return 0;
    if (r >= 255.0) return 255;
    return (uint8_t)r;
}
struct FMaxOp {
    __device__ __forceinline__ double operator()(double a, double b) const { return fmax(a, b); }
};
__global__ void linear_max_kernel(const double* xyza, size_t n, double exposure, double* comp_max) {
    // max over pixels of lin*exposure for r,g,b: fold(0.0, f64::max); block reduce + atomic on the ordered bits
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    double m = 0.0;
    if (i < n) {
        double lin[3];
        xyz_to_linear(xyza + 4 * i, lin);
        m = fmax(fmax(fmax(0.0, lin[0] * exposure), lin[1] * exposure), lin[2] * exposure);
    }
    typedef cub::BlockReduce<double, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    double bm = BR(tmp).Reduce(m, FMaxOp());
    if (threadIdx.x == 0 && bm > 0.0)  // non-negative doubles order like their bit patterns
        atomicMax(reinterpret_cast<unsigned long long*>(comp_max), (unsigned long long)__double_as_longlong(bm));
}
__global__ void tonemap_kernel(const double* xyza, size_t n, int method, double exposure, const double* comp_max,
                               uint8_t* rgb) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double c[3];
    xyz_to_linear(xyza + 4 * i, c);
    c[0] *= exposure; c[1] *= exposure; c[2] *= exposure;
    if (method == GRRT_TONEMAP_REINHARD) {
        double l_in = (0.2126 * c[0] + 0.7152 * c[1]) + 0.0722 * c[2];
        if (l_in > 0.0) {
            double l_out = l_in / (1.0 + l_in);
            double f = l_out / l_in;
            c[0] *= f; c[1] *= f; c[2] *= f;
        }
    } else {
        double mc = *comp_max;
        double scale = mc > 0.0 ? 1.0 / mc : 1.0;
        c[0] = scale * c[0]; c[1] = scale * c[1]; c[2] = scale * c[2];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) rgb[3 * i + k] = to_u8(compand_srgb(fmax(c[k], 0.0)) * 255.0);
}

// ---- FP64 issue-rate microbenchmark -----------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_chain_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0, x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0,
           x7 = x0 + 7.0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// mean over the valid sub-samples, summed in stratum order (raytracer.rs:341-379)
struct PeerFrames {
    unsigned n;
    double* xyza[GRRT_MAX_PEERS];
};
__global__ void grrt_exact_resolve(const double* sub_xyza, const uint8_t* sub_valid, const uint32_t* items,
                                   unsigned long long item_offset, unsigned long long item_stride,
                                   unsigned long long n_items, unsigned ss, double* xyza, PeerFrames peers) {
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    const unsigned long long pid = items[item_offset + i * item_stride];
    double sx = 0.0, sy = 0.0, sz = 0.0, sa = 0.0;
    unsigned valid = 0;
    for (unsigned k = 0; k < ss; ++k) {
        unsigned long long u = i * ss + k;
        if (sub_valid[u]) {
            const double* c = sub_xyza + 4 * u;
            sx = sx + c[0]; sy = sy + c[1]; sz = sz + c[2]; sa = sa + c[3];
            ++valid;
        }
    }
    if (valid > 0) {
        double inv = 1.0 / (double)valid;
        const double m0 = sx * inv, m1 = sy * inv, m2 = sz * inv, m3 = sa * inv;
        if (peers.n == 0) {
            double* dst = xyza + 4ull * pid;
            dst[0] = m0; dst[1] = m1; dst[2] = m2; dst[3] = m3;
        } else {
            for (unsigned p = 0; p < peers.n; ++p) {  // the mean goes into every rank's frame
                double* dst = peers.xyza[p] + 4ull * pid;
                dst[0] = m0; dst[1] = m1; dst[2] = m2; dst[3] = m3;
            }
        }
    }
}

int launch_trace(GrrtCtx* ctx, const GrrtDeviceScene* ds, const TraceArgs& a, cudaStream_t st) {
    TraceLaunch l{ctx->num_sms, st};
    // Kerr-Schild always runs the exact (no-FMA) kernels; the smooth geometries use the FMA build unless disabled
    const bool fast = ctx->use_fma && ds->s.geom != GRRT_GEOM_KERR;
    int rc = fast ? grrt_fast::launch_trace(ds->s, a, ds->has_vol, l) : grrt_exact::launch_trace(ds->s, a, ds->has_vol, l);
    if (rc != 0) {
        ctx->last_error = std::string("trace_kernel launch: ") + cudaGetErrorString((cudaError_t)rc);
        return GRRT_ERR_CUDA;
    }
    ++ctx->launches;
    return GRRT_OK;
}

int ensure_cub_temp(GrrtCtx* ctx, size_t bytes) {
    if (bytes <= ctx->cub_temp_bytes) return GRRT_OK;
    if (ctx->d_cub_temp) cudaFree(ctx->d_cub_temp);
    ctx->d_cub_temp = nullptr;
    ctx->cub_temp_bytes = 0;
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_cub_temp, bytes));
    ctx->cub_temp_bytes = bytes;
    return GRRT_OK;
}

// RAII device allocation from the stream-ordered pool: after warm-up neither the allocation nor the release
// touches the driver's page tables or synchronises the device (plain cudaMalloc/cudaFree cost up to hundreds
// of milliseconds per frame here and serialise against running kernels).
struct DevBuf {
    void* p = nullptr;
    cudaStream_t st = nullptr;
    ~DevBuf() { if (p) cudaFreeAsync(p, st); }
    cudaError_t alloc(size_t bytes, cudaStream_t stream) {
        st = stream;
        return cudaMallocAsync(&p, bytes ? bytes : 1, stream);
    }
    template <class T> T* as() { return static_cast<T*>(p); }
};

}  // namespace

static void preload_small_kernels(GrrtCtx* ctx);

// =========================================================================================================
extern "C" {

const char* grrt_version(void) { return "grrt-b200 0.1 (sm_100a)"; }

static int init_one(int device_id, GrrtCtx** out) {
    if (!out) return GRRT_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0 || device_id < 0 || device_id >= count) return GRRT_ERR_CUDA;
    GrrtCtx* ctx = new GrrtCtx();
    ctx->device = device_id;
    if (cudaSetDevice(device_id) != cudaSuccess) { delete ctx; return GRRT_ERR_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) { delete ctx; return GRRT_ERR_CUDA; }
    if (prop.major < 10) {  // the only code in this library is sm_100a SASS
        delete ctx;
        return GRRT_ERR_CUDA;
    }
    ctx->num_sms = prop.multiProcessorCount;
    if (const char* e = std::getenv("GRRT_FMA")) ctx->use_fma = !(e[0] == '0');
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return GRRT_ERR_CUDA; }
    {  // keep freed blocks cached in the device's default memory pool instead of returning them to the OS
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device_id) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    double lut[256];
    for (int i = 0; i < 256; ++i) lut[i] = inv_compand_srgb_host((double)i / 255.0);
    if (cudaMalloc(&ctx->d_srgb_lut, sizeof lut) != cudaSuccess ||
        cudaMemcpy(ctx->d_srgb_lut, lut, sizeof lut, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMalloc(&ctx->d_queue, sizeof(unsigned long long)) != cudaSuccess ||
        cudaMalloc(&ctx->d_counters, 16 * sizeof(unsigned long long)) != cudaSuccess) {
        grrt_destroy(ctx);
        return GRRT_ERR_CUDA;
    }
    // CUDA loads a kernel's code on its first launch on a device, and that load waits for the kernels already running
    // there: with several devices driven from one thread the first frame would trace its passes one device after the
    // other.  Launch the small kernels once with no work now.
    preload_small_kernels(ctx);
    *out = ctx;
    return GRRT_OK;
}

int grrt_init(int device_id, GrrtCtx** out) { return init_one(device_id, out); }

int grrt_init_multi(int n_devices, const int* device_ids, GrrtCtx** out) {
    if (!out) return GRRT_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (n_devices < 1 || n_devices > GRRT_MAX_PEERS || !device_ids) return GRRT_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < n_devices; ++i)
        for (int j = 0; j < i; ++j)
            if (device_ids[i] == device_ids[j]) return GRRT_ERR_INVALID_ARGUMENT;
    GrrtCtx* primary = nullptr;
    int rc = init_one(device_ids[0], &primary);
    if (rc) return rc;
    for (int i = 1; i < n_devices; ++i) {
        GrrtCtx* m = nullptr;
        rc = init_one(device_ids[i], &m);
        if (rc == GRRT_OK) {
            // the member's kernels store finished pixels into the primary device's frame over NVLink
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, device_ids[i], device_ids[0]) != cudaSuccess || !can) rc = GRRT_ERR_CUDA;
        }
        if (rc == GRRT_OK) {
            cudaSetDevice(device_ids[i]);
            cudaError_t e = cudaDeviceEnablePeerAccess(device_ids[0], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
            if (e != cudaSuccess) rc = GRRT_ERR_CUDA;
        }
        if (rc == GRRT_OK) {  // frame buffers come from the primary's stream-ordered pool: open it to the member
            cudaMemPool_t pool;
            cudaMemAccessDesc desc;
            std::memset(&desc, 0, sizeof desc);
            desc.location.type = cudaMemLocationTypeDevice;
            desc.location.id = device_ids[i];
            desc.flags = cudaMemAccessFlagsProtReadWrite;
            if (cudaDeviceGetDefaultMemPool(&pool, device_ids[0]) != cudaSuccess || cudaMemPoolSetAccess(pool, &desc, 1) != cudaSuccess)
                rc = GRRT_ERR_CUDA;
        }
        if (rc != GRRT_OK) {
            if (m) grrt_destroy(m);
            grrt_destroy(primary);
            return rc;
        }
        m->use_fma = primary->use_fma;
        primary->members.push_back(m);
    }
    cudaSetDevice(device_ids[0]);
    *out = primary;
    return GRRT_OK;
}

int grrt_device_count(const GrrtCtx* ctx) { return ctx ? 1 + (int)ctx->members.size() : 0; }

void grrt_destroy(GrrtCtx* ctx) {
    if (!ctx) return;
    for (size_t k = 0; k < ctx->d_part_items.size(); ++k) {
        if (!ctx->d_part_items[k]) continue;
        cudaSetDevice(k == 0 ? ctx->device : ctx->members[k - 1]->device);
        cudaFree(ctx->d_part_items[k]);
    }
    for (GrrtCtx* m : ctx->members) grrt_destroy(m);
    ctx->members.clear();
    cudaSetDevice(ctx->device);
    if (ctx->d_srgb_lut) cudaFree(ctx->d_srgb_lut);
    if (ctx->d_queue) cudaFree(ctx->d_queue);
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    if (ctx->d_cub_temp) cudaFree(ctx->d_cub_temp);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    delete ctx;
}

int grrt_set_option(GrrtCtx* ctx, const char* name, int value) {
    if (!ctx || !name) return GRRT_ERR_INVALID_ARGUMENT;
    if (std::strcmp(name, "fma") == 0) {
        ctx->use_fma = value != 0;
        for (GrrtCtx* m : ctx->members) m->use_fma = ctx->use_fma;
        return GRRT_OK;
    }
    return GRRT_ERR_INVALID_ARGUMENT;
}

unsigned long long grrt_kernel_launch_count(const GrrtCtx* ctx) {
    if (!ctx) return 0;
    unsigned long long n = ctx->launches;
    for (const GrrtCtx* m : ctx->members) n += m->launches;
    return n;
}

const char* grrt_last_error(const GrrtCtx* ctx) { return ctx ? ctx->last_error.c_str() : "no context"; }

namespace {
void set_camera(DScene& s, const GrrtCamera& cam) {
    for (int i = 0; i < 4; ++i) {
        s.cam_pos[i] = cam.position[i];
        s.cam_vel[i] = cam.velocity[i];
        for (int j = 0; j < 4; ++j) s.tet[i][j] = cam.tetrad[i][j];
    }
    s.cam_rows = (double)cam.rows;
    s.cam_cols = (double)cam.cols;
    s.cam_scale = 2.0 * std::tan(cam.alpha / 2.0) / (double)cam.rows;  // camera.rs:217-224
    s.cam_hand = cam.spatial_handedness;
    s.cam_sig = cam.spatial_signature;
}
}  // namespace

int grrt_scene_set_camera(GrrtCtx* ctx, GrrtDeviceScene* ds, const GrrtCamera* camera) {
    if (!ctx || !ds || !camera) return GRRT_ERR_INVALID_ARGUMENT;
    if (camera->rows <= 0 || camera->cols <= 0) return fail(ctx, GRRT_ERR_INVALID_ARGUMENT, "camera has no pixels");
    set_camera(ds->s, *camera);  // the scene block travels as a kernel argument: nothing to copy to the device
    for (GrrtDeviceScene* r : ds->replicas) set_camera(r->s, *camera);
    return GRRT_OK;
}

static int scene_upload_one(GrrtCtx* ctx, const GrrtScene* sc, GrrtDeviceScene** out);

// The scene of the primary device copied to a member device: the tables travel device to device over NVLink (one
// host -> device transfer per frame, not one per GPU), the parameter block is patched to the member's copies.
static int scene_replicate(GrrtCtx* ctx, const GrrtDeviceScene* src, GrrtCtx* m, GrrtDeviceScene** out) {
    *out = nullptr;
    GrrtDeviceScene* r = new GrrtDeviceScene();
    r->s = src->s;
    r->adaptive = src->adaptive;
    r->has_vol = src->has_vol;
    auto bail = [&](GrrtCtx* c, const char* what) {
        m->last_error = std::string(what) + ": " + cudaGetErrorString(cudaGetLastError());
        grrt_scene_free(c, r);
        return GRRT_ERR_CUDA;
    };
    cudaEvent_t uploaded;   // the primary's host -> device copies are complete
    if (cudaSetDevice(ctx->device) != cudaSuccess || cudaEventCreateWithFlags(&uploaded, cudaEventDisableTiming) != cudaSuccess)
        return bail(m, "cudaEventCreate");
    cudaEventRecord(uploaded, ctx->stream);
    if (cudaSetDevice(m->device) != cudaSuccess) { cudaEventDestroy(uploaded); return bail(m, "cudaSetDevice"); }
    cudaStreamWaitEvent(m->stream, uploaded, 0);
    std::vector<void*> copy(src->allocs.size(), nullptr);
    for (size_t k = 0; k < src->allocs.size(); ++k) {
        void* p = nullptr;
        if (cudaMallocAsync(&p, src->alloc_bytes[k], m->stream) != cudaSuccess) { cudaEventDestroy(uploaded); return bail(m, "cudaMallocAsync"); }
        r->allocs.push_back(p);
        r->alloc_bytes.push_back(src->alloc_bytes[k]);
        copy[k] = p;
        if (cudaMemcpyPeerAsync(p, m->device, src->allocs[k], ctx->device, src->alloc_bytes[k], m->stream) != cudaSuccess) {
            cudaEventDestroy(uploaded);
            return bail(m, "cudaMemcpyPeerAsync");
        }
    }
    cudaSetDevice(ctx->device);
    cudaEventDestroy(uploaded);   // released once the recorded work has completed
    auto patch = [&](const void* p) -> const void* {
        if (!p) return p;
        for (size_t k = 0; k < src->allocs.size(); ++k)
            if (p == src->allocs[k]) return copy[k];
        return p;
    };
    DScene& s = r->s;
    s.srgb_lut = m->d_srgb_lut;
    s.celestial.rgba = static_cast<const uint8_t*>(patch(s.celestial.rgba));
    s.bb_log_t = static_cast<const double*>(patch(s.bb_log_t));
    s.bb_xyz = static_cast<const double*>(patch(s.bb_xyz));
    for (int k = 0; k < GRRT_MAX_OBJECTS; ++k) {
        DObject& o = s.obj[k];
        o.tex.rgba = static_cast<const uint8_t*>(patch(o.tex.rgba));
        o.temp.lut_r = static_cast<const double*>(patch(o.temp.lut_r));
        o.temp.lut_t = static_cast<const double*>(patch(o.temp.lut_t));
        o.perm = static_cast<const uint8_t*>(patch(o.perm));
    }
    *out = r;
    return GRRT_OK;
}

int grrt_scene_upload(GrrtCtx* ctx, const GrrtScene* sc, GrrtDeviceScene** out) {
    if (!ctx || !sc || !out) return GRRT_ERR_INVALID_ARGUMENT;
    int rc = scene_upload_one(ctx, sc, out);
    if (rc) return rc;
    for (GrrtCtx* m : ctx->members) {  // the same scene on every device of a multi-device context
        GrrtDeviceScene* r = nullptr;
        rc = scene_replicate(ctx, *out, m, &r);
        if (rc) {
            ctx->last_error = m->last_error;
            grrt_scene_free(ctx, *out);
            *out = nullptr;
            cudaSetDevice(ctx->device);
            return rc;
        }
        (*out)->replicas.push_back(r);
    }
    cudaSetDevice(ctx->device);
    return GRRT_OK;
}

static int scene_upload_one(GrrtCtx* ctx, const GrrtScene* sc, GrrtDeviceScene** out) {
    if (!ctx || !sc || !out) return GRRT_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (sc->n_objects > GRRT_MAX_OBJECTS) return fail(ctx, GRRT_ERR_INVALID_ARGUMENT, "too many objects");
    if (sc->camera.rows <= 0 || sc->camera.cols <= 0) return fail(ctx, GRRT_ERR_INVALID_ARGUMENT, "camera has no pixels");
    if (sc->kerr_rhs_mode != GRRT_KERR_RHS_REFERENCE_FD)
        return fail(ctx, GRRT_ERR_INVALID_CONFIGURATION, "kerr_rhs_mode: only the reference's finite-difference right-hand side exists");
    GrrtDeviceScene* ds = new GrrtDeviceScene();
    auto bail = [&](int rc) {
        grrt_scene_free(ctx, ds);
        return rc;
    };
    DScene& s = ds->s;
    std::memset(&s, 0, sizeof s);
    ds->adaptive = sc->adaptive;
    s.geom = sc->geometry.kind;
    s.kerr_mode = sc->kerr_rhs_mode;
    s.radius = sc->geometry.radius;
    s.a = sc->geometry.a;
    s.heps = sc->geometry.horizon_epsilon;
    {  // kerr.rs:382-394 / kerr_bl.rs:482-492
        double m = (s.geom == GRRT_GEOM_KERR) ? 0.5 * s.radius : s.radius / 2.0;
        s.horizon_enabled = !(std::fabs(s.a) > s.radius / 2.0);
        double disc = std::fmax(m * m - s.a * s.a, 0.0);
        double rp = m + std::sqrt(disc);
        s.r_plus_eps = rp + s.heps;
    }
    s.n_objects = (int)sc->n_objects;
    set_camera(s, sc->camera);
    s.max_steps = sc->integration.max_steps;
    s.max_radius_sq = sc->integration.max_radius * sc->integration.max_radius;  // integrator.rs:62
    s.step_size = sc->integration.step_size;
    s.epsilon = sc->integration.epsilon;
    s.inv_epsilon = 1.0 / sc->integration.epsilon;
    s.hit_opacity_threshold = sc->adaptive.object_hit_opacity_threshold;
    s.celestial_temperature = sc->celestial_temperature;
    s.srgb_lut = ctx->d_srgb_lut;
    std::map<const uint8_t*, const uint8_t*> tex_cache;
    int rc = convert_texture(ctx, ds, sc->celestial, &s.celestial, tex_cache);
    if (rc) return bail(rc);
    bool need_bb = sc->celestial.kind == GRRT_TEX_BLACKBODY;
    int n_vol = 0;
    for (uint32_t k = 0; k < sc->n_objects; ++k) {
        const GrrtObject& o = sc->objects[k];
        DObject& d = s.obj[k];
        d.kind = o.kind;
        rc = convert_texture(ctx, ds, o.texture, &d.tex, tex_cache);
        if (rc) return bail(rc);
        need_bb = need_bb || o.texture.kind == GRRT_TEX_BLACKBODY;
        d.temp.kind = o.temperature.kind;
        d.temp.n = o.temperature.n;
        d.temp.constant = o.temperature.constant;
        d.temp.r_isco = o.temperature.r_isco;
        if (o.kind != GRRT_OBJ_SPHERE && o.temperature.kind == GRRT_TEMP_LUT) {
            if ((rc = upload<double>(ctx, ds, o.temperature.lut_r, o.temperature.n, &d.temp.lut_r))) return bail(rc);
            if ((rc = upload<double>(ctx, ds, o.temperature.lut_t, o.temperature.n, &d.temp.lut_t))) return bail(rc);
        }
        d.radius = o.radius;
        d.center_norm = std::sqrt(o.position[0] * o.position[0] + o.position[1] * o.position[1] + o.position[2] * o.position[2]);
        d.axis_is_z = (o.axis[0] == 0.0 && o.axis[1] == 0.0 && o.axis[2] == 1.0) ? 1 : 0;
        // summary for segment_needs_cartesian() (grrt_device.cuh)
        if (o.kind == GRRT_OBJ_DISC) {
            s.pre_flags |= PRE_HAS_DISC;
        } else if (o.kind == GRRT_OBJ_SPHERE) {
            const double Rs = o.radius * (1.0 + 1e-9) + 1e-300;
            const double far = d.center_norm + Rs, near = d.center_norm - Rs;
            s.pre_sph_far2[s.pre_n_spheres] = far * far * (1.0 + 1e-12);
            s.pre_sph_near2[s.pre_n_spheres] = near > 0.0 ? near * near * (1.0 - 1e-12) : -1.0;
            ++s.pre_n_spheres;
        } else {
            s.pre_flags |= d.axis_is_z ? PRE_HAS_VOL : PRE_ALWAYS;
            s.pre_vol_h = std::max(s.pre_vol_h, 3.0 * o.thickness * (1.0 + 1e-9));
        }
        d.sphere_temperature = o.sphere_temperature;
        d.r_in = o.inner_radius;
        d.r_out = o.outer_radius;
        for (int i = 0; i < 3; ++i) {
            d.pos[i] = o.position[i];
            d.axis[i] = o.axis[i];
            d.e1[i] = o.e1[i];
            d.e2[i] = o.e2[i];
            d.noise_scale[i] = o.noise_scale[i];
        }
        d.num_octaves = o.num_octaves;
        d.vol_max_steps = o.vol_max_steps;
        d.vol_step = o.vol_step_size;
        d.thickness = o.thickness;
        d.density_multiplier = o.density_multiplier;
        d.bref_temperature = o.brightness_reference_temperature;
        d.absorption = o.absorption;
        d.scattering = o.scattering;
        d.noise_offset = o.noise_offset;
        if (o.kind == GRRT_OBJ_VOLUMETRIC_DISC) {
            d.perm_slot = n_vol++;
            if ((rc = upload<uint8_t>(ctx, ds, o.perlin_perm, 256, &d.perm))) return bail(rc);
        }
    }
    s.n_vol = (uint32_t)n_vol;
    ds->has_vol = n_vol > 0;
    if (need_bb) {
        if (!sc->bb_lut_log_t || !sc->bb_lut_xyz || sc->bb_lut_n == 0)
            return bail(fail(ctx, GRRT_ERR_INVALID_ARGUMENT, "BlackBody texture without a black-body LUT"));
        s.bb_n = sc->bb_lut_n;
        if ((rc = upload<double>(ctx, ds, sc->bb_lut_log_t, sc->bb_lut_n, &s.bb_log_t))) return bail(rc);
        if ((rc = upload<double>(ctx, ds, sc->bb_lut_xyz, (size_t)sc->bb_lut_n * 3, &s.bb_xyz))) return bail(rc);
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *out = ds;
    return GRRT_OK;
}

void grrt_scene_free(GrrtCtx* ctx, GrrtDeviceScene* ds) {
    if (!ds) return;
    for (size_t k = 0; k < ds->replicas.size(); ++k)
        grrt_scene_free(ctx && k < ctx->members.size() ? ctx->members[k] : nullptr, ds->replicas[k]);
    ds->replicas.clear();
    if (ctx) cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();  // the scene may be in use on a caller's stream
    for (void* p : ds->allocs) cudaFreeAsync(p, ctx ? ctx->stream : nullptr);
    delete ds;
}

namespace {
struct TraceOpts {   // everything beyond the plain item list of grrt_trace_items_device
    uint64_t item_offset = 0, item_stride = 1;   // item i of this call is d_items[item_offset + i * item_stride]
    uint32_t n_peers = 0;                        // > 0: results go at their pixel id into every peer's frame
    double* const* xyza_peers = nullptr;
    uint8_t* const* class_peers = nullptr;
    uint8_t* const* stop_peers = nullptr;
    uint32_t* const* steps_peers = nullptr;      // nullable, and so is every entry
};

int trace_items_impl(GrrtCtx* ctx, const GrrtDeviceScene* ds, const uint32_t* d_items, uint64_t first_item,
                     uint64_t n_items, uint32_t from_row, uint32_t from_col, uint32_t to_row, uint32_t to_col,
                     uint32_t samples_per_axis, double* d_xyza, uint8_t* d_class, uint8_t* d_stop, uint32_t* d_steps,
                     uint64_t* d_counters, void* cuda_stream, const TraceOpts& o) {
    if (to_row < from_row || to_col < from_col) return fail(ctx, GRRT_ERR_INVALID_ARGUMENT, "empty or inverted section");
    if (o.item_stride == 0) return fail(ctx, GRRT_ERR_INVALID_ARGUMENT, "item_stride must be at least 1");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    if (n_items == 0) return GRRT_OK;
    const uint32_t n_peers = o.n_peers;
    TraceArgs a;
    std::memset(&a, 0, sizeof a);
    a.items = d_items;
    a.item_offset = o.item_offset;
    a.item_stride = o.item_stride;
    a.first_item = first_item;
    a.n_items = n_items;
    a.from_row = from_row;
    a.from_col = from_col;
    a.width = to_col - from_col;
    a.height = to_row - from_row;
    a.queue = ctx->d_queue;
    a.counters = (unsigned long long*)d_counters;
    PeerFrames pf;
    std::memset(&pf, 0, sizeof pf);
    pf.n = n_peers;
    for (uint32_t p = 0; p < n_peers; ++p) pf.xyza[p] = o.xyza_peers[p];
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_queue, 0, sizeof(unsigned long long), st));
    if (samples_per_axis == 0) {
        a.spa = 0;
        a.xyza = d_xyza;
        a.cls = d_class;
        a.stop = d_stop;
        a.steps = d_steps;
        a.n_peers = n_peers;
        for (uint32_t p = 0; p < n_peers; ++p) {
            a.xyza_p[p] = o.xyza_peers[p];
            a.cls_p[p] = o.class_peers[p];
            a.stop_p[p] = o.stop_peers[p];
            a.steps_p[p] = o.steps_peers ? o.steps_peers[p] : nullptr;
        }
        const bool whole = !d_items && first_item == 0 && n_items == (uint64_t)a.width * a.height;
        if (whole) {  // enumerate 8x4 pixel tiles so that a warp starts on a compact patch of the image
            a.tiled = 1;
            a.tiles_x = (a.width + 7) / 8;
            a.n_units = (unsigned long long)a.tiles_x * ((a.height + 3) / 4) * 32ull;
        } else {
            a.n_units = n_items;
        }
        return launch_trace(ctx, ds, a, st);
    }
    if (!d_items) return fail(ctx, GRRT_ERR_INVALID_ARGUMENT, "sub-sample mode needs an explicit item list");
    // sub-sample mode: trace N*N rays per item into scratch, then average in stratum order
    const unsigned ss = samples_per_axis * samples_per_axis;
    const uint64_t chunk_items = std::max<uint64_t>(1, (uint64_t)(1ull << 28) / ss);  // <= 8 GiB of scratch
    DevBuf sub, valid;
    const uint64_t cap = std::min<uint64_t>(chunk_items, n_items);
    CUDA_TRY(ctx, sub.alloc(cap * ss * 4 * sizeof(double), st));
    CUDA_TRY(ctx, valid.alloc(cap * ss, st));
    for (uint64_t off = 0; off < n_items; off += cap) {
        uint64_t m = std::min<uint64_t>(cap, n_items - off);
        a.item_offset = o.item_offset + off * o.item_stride;
        a.n_items = m;
        a.spa = samples_per_axis;
        a.n_units = m * ss;
        a.xyza = sub.as<double>();
        a.sub_valid = valid.as<uint8_t>();
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_queue, 0, sizeof(unsigned long long), st));
        int rc = launch_trace(ctx, ds, a, st);
        if (rc) return rc;
        grrt_exact_resolve<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(sub.as<double>(), valid.as<uint8_t>(), d_items,
                                                                                a.item_offset, a.item_stride, m, ss, d_xyza, pf);
        ++ctx->launches;
        CUDA_TRY(ctx, cudaGetLastError());
    }
    return GRRT_OK;  // the scratch buffers are freed in stream order
}
}  // namespace

int grrt_trace_items_device(GrrtCtx* ctx, const GrrtDeviceScene* ds, const uint32_t* d_items, uint64_t first_item,
                            uint64_t n_items, uint32_t from_row, uint32_t from_col, uint32_t to_row, uint32_t to_col,
                            uint32_t samples_per_axis, double* d_xyza, uint8_t* d_class, uint8_t* d_stop,
                            uint32_t* d_steps, uint64_t* d_counters, void* cuda_stream) {
    if (!ctx || !ds || !d_xyza) return GRRT_ERR_INVALID_ARGUMENT;
    return trace_items_impl(ctx, ds, d_items, first_item, n_items, from_row, from_col, to_row, to_col, samples_per_axis, d_xyza,
                            d_class, d_stop, d_steps, d_counters, cuda_stream, TraceOpts());
}

int grrt_trace_items_device_peers(GrrtCtx* ctx, const GrrtDeviceScene* ds, const uint32_t* d_items, uint64_t item_offset,
                                  uint64_t item_stride, uint64_t n_items, uint32_t from_row, uint32_t from_col,
                                  uint32_t to_row, uint32_t to_col, uint32_t samples_per_axis, uint32_t n_peers,
                                  double* const* d_xyza_peers, uint8_t* const* d_class_peers, uint8_t* const* d_stop_peers,
                                  uint32_t* const* d_steps_peers, uint64_t* d_counters, void* cuda_stream) {
    if (!ctx || !ds || !d_items || !d_xyza_peers || !d_class_peers || !d_stop_peers) return GRRT_ERR_INVALID_ARGUMENT;
    if (n_peers == 0 || n_peers > GRRT_MAX_PEERS) return fail(ctx, GRRT_ERR_INVALID_ARGUMENT, "1..GRRT_MAX_PEERS peers");
    for (uint32_t p = 0; p < n_peers; ++p)
        if (!d_xyza_peers[p] || !d_class_peers[p] || !d_stop_peers[p]) return fail(ctx, GRRT_ERR_INVALID_ARGUMENT, "null peer buffer");
    TraceOpts o;
    o.item_offset = item_offset;
    o.item_stride = item_stride;
    o.n_peers = n_peers;
    o.xyza_peers = d_xyza_peers;
    o.class_peers = d_class_peers;
    o.stop_peers = d_stop_peers;
    o.steps_peers = d_steps_peers;
    return trace_items_impl(ctx, ds, d_items, 0, n_items, from_row, from_col, to_row, to_col, samples_per_axis, d_xyza_peers[0],
                            nullptr, nullptr, nullptr, d_counters, cuda_stream, o);
}

int grrt_ipc_alloc(GrrtCtx* ctx, size_t bytes, void** d_ptr, unsigned char handle_out[64]) {
    if (!ctx || !d_ptr || !handle_out) return GRRT_ERR_INVALID_ARGUMENT;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    void* p = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&p, bytes ? bytes : 1));   // plain cudaMalloc: pool allocations cannot be exported
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        CUDA_TRY(ctx, e);
    }
    std::memcpy(handle_out, &h, 64);
    *d_ptr = p;
    return GRRT_OK;
}
int grrt_ipc_open(GrrtCtx* ctx, const unsigned char handle[64], void** d_ptr) {
    if (!ctx || !handle || !d_ptr) return GRRT_ERR_INVALID_ARGUMENT;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    CUDA_TRY(ctx, cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return GRRT_OK;
}
int grrt_ipc_close(GrrtCtx* ctx, void* d_ptr) {
    if (!ctx || !d_ptr) return GRRT_ERR_INVALID_ARGUMENT;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaIpcCloseMemHandle(d_ptr));
    return GRRT_OK;
}
int grrt_ipc_free(GrrtCtx* ctx, void* d_ptr) {
    if (!ctx || !d_ptr) return GRRT_ERR_INVALID_ARGUMENT;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaFree(d_ptr));
    return GRRT_OK;
}

int grrt_select_mask_device(GrrtCtx* ctx, const GrrtDeviceScene* ds, const double* d_xyza, const uint8_t* d_class,
                            uint32_t rows, uint32_t cols, uint32_t* d_selected, uint32_t* d_count,
                            double* minimum_luminance_out, void* cuda_stream) {
    if (!ctx || !ds || !d_xyza || !d_class || !d_selected || !d_count) return GRRT_ERR_INVALID_ARGUMENT;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    const size_t n = (size_t)rows * cols;
    if (n == 0) {
        CUDA_TRY(ctx, cudaMemsetAsync(d_count, 0, sizeof(uint32_t), st));
        if (minimum_luminance_out) *minimum_luminance_out = 0.0;
        return GRRT_OK;
    }
    DevBuf keys_in, keys_out, flags, minlum;
    CUDA_TRY(ctx, flags.alloc(n, st));
    CUDA_TRY(ctx, minlum.alloc(sizeof(double), st));
    const unsigned long long* key_at = nullptr;
    if (!ds->adaptive.has_minimum_luminance) {  // 99th percentile by a full radix sort of total_cmp keys
        CUDA_TRY(ctx, keys_in.alloc(n * 8, st));
        CUDA_TRY(ctx, keys_out.alloc(n * 8, st));
        luminance_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_xyza, n, keys_in.as<unsigned long long>());
        ++ctx->launches;
        size_t tmp = 0;
        CUDA_TRY(ctx, cub::DeviceRadixSort::SortKeys(nullptr, tmp, keys_in.as<unsigned long long>(),
                                                     keys_out.as<unsigned long long>(), (long long)n, 0, 64, st));
        int rc = ensure_cub_temp(ctx, tmp);
        if (rc) return rc;
        CUDA_TRY(ctx, cub::DeviceRadixSort::SortKeys(ctx->d_cub_temp, tmp, keys_in.as<unsigned long long>(),
                                                     keys_out.as<unsigned long long>(), (long long)n, 0, 64, st));
        size_t index = (size_t)((double)(n - 1) * 0.99);  // raytracer.rs:126
        key_at = keys_out.as<unsigned long long>() + index;
    }
    mask_flags_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_xyza, d_class, rows, cols, ds->adaptive, key_at,
                                                                    minlum.as<double>(), flags.as<uint8_t>());
    ++ctx->launches;
    CUDA_TRY(ctx, cudaGetLastError());
    {
        cub::CountingInputIterator<uint32_t> ids(0);
        size_t tmp = 0;
        CUDA_TRY(ctx, cub::DeviceSelect::Flagged(nullptr, tmp, ids, flags.as<uint8_t>(), d_selected, d_count, (long long)n, st));
        int rc = ensure_cub_temp(ctx, tmp);
        if (rc) return rc;
        CUDA_TRY(ctx, cub::DeviceSelect::Flagged(ctx->d_cub_temp, tmp, ids, flags.as<uint8_t>(), d_selected, d_count,
                                                 (long long)n, st));
    }
    double ml = 0.0;
    CUDA_TRY(ctx, cudaMemcpyAsync(&ml, minlum.p, sizeof ml, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    if (minimum_luminance_out) *minimum_luminance_out = ml;
    return GRRT_OK;
}

int grrt_tonemap_device(GrrtCtx* ctx, const double* d_xyza, size_t n, int method, double exposure, uint8_t* d_rgb,
                        void* cuda_stream) {
    if (!ctx || !d_xyza || !d_rgb) return GRRT_ERR_INVALID_ARGUMENT;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    if (n == 0) return GRRT_OK;
    double* d_max = reinterpret_cast<double*>(ctx->d_counters + 15);
    if (method == GRRT_TONEMAP_GLOBAL_LINEAR) {
        CUDA_TRY(ctx, cudaMemsetAsync(d_max, 0, sizeof(double), st));
        linear_max_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_xyza, n, exposure, d_max);
        ++ctx->launches;
    }
    tonemap_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_xyza, n, method, exposure, d_max, d_rgb);
    ++ctx->launches;
    CUDA_TRY(ctx, cudaGetLastError());
    return GRRT_OK;
}

int grrt_tonemap(GrrtCtx* ctx, const double* xyza, size_t n, int method, double exposure, uint8_t* rgb_out) {
    if (!ctx || !xyza || !rgb_out) return GRRT_ERR_INVALID_ARGUMENT;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    DevBuf in, out;
    CUDA_TRY(ctx, in.alloc(n * 32, ctx->stream));
    CUDA_TRY(ctx, out.alloc(n * 3, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(in.p, xyza, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    int rc = grrt_tonemap_device(ctx, in.as<double>(), n, method, exposure, out.as<uint8_t>(), ctx->stream);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(rgb_out, out.p, n * 3, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return GRRT_OK;
}

int grrt_render(GrrtCtx* ctx, const GrrtScene* scene, uint32_t from_row, uint32_t from_col, uint32_t to_row,
                uint32_t to_col, double* xyza_out, uint8_t* ray_class_out, uint8_t* stop_reason_out, uint32_t* steps_out,
                GrrtStats* stats) {
    if (!ctx || !scene || !xyza_out) return GRRT_ERR_INVALID_ARGUMENT;
    if (to_row < from_row || to_col < from_col) return fail(ctx, GRRT_ERR_INVALID_ARGUMENT, "inverted section");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    GrrtDeviceScene* ds = nullptr;
    int rc = grrt_scene_upload(ctx, scene, &ds);
    if (rc) return rc;
    rc = grrt_render_resident(ctx, ds, from_row, from_col, to_row, to_col, xyza_out, ray_class_out, stop_reason_out, steps_out,
                              stats);
    grrt_scene_free(ctx, ds);
    return rc;
}

namespace {
// ---- supersampling order --------------------------------------------------------------------------------
// The base pass is the pilot of the second pass: a flagged pixel's 16 sub-sample rays take about as many RKF45 steps
// as its centre ray did.  The flagged pixels are therefore traced longest first (the work queue then ends on the
// short rays, not on a straggler) and, on several devices, dealt round-robin in that order so that every device gets
// the same share of the expensive ones.  The image does not depend on the order: every pixel's mean is taken over
// its own strata in stratum order.
__global__ void cost_keys_kernel(const uint32_t* sel, uint32_t n, const uint32_t* steps, uint32_t* keys) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = ~steps[sel[i]];   // ascending ~steps = descending steps; the radix sort is stable
}
// the pixels the second pass rewrote, packed for the read-back: [n] ids then [n][4] values
__global__ void gather_pixels_kernel(const uint32_t* sel, uint32_t n, const double* xyza, double* packed) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* src = xyza + 4ull * sel[i];
    double* dst = packed + 4ull * i;
    dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
}
}  // namespace

int grrt_order_by_cost_device(GrrtCtx* ctx, const uint32_t* d_selected, uint32_t n_selected, const uint32_t* d_steps,
                              uint32_t* d_ordered, void* cuda_stream) {
    if (!ctx || !d_selected || !d_steps || !d_ordered) return GRRT_ERR_INVALID_ARGUMENT;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    if (n_selected == 0) return GRRT_OK;
    DevBuf k_in, k_out;
    CUDA_TRY(ctx, k_in.alloc((size_t)n_selected * 4, st));
    CUDA_TRY(ctx, k_out.alloc((size_t)n_selected * 4, st));
    cost_keys_kernel<<<(n_selected + 255) / 256, 256, 0, st>>>(d_selected, n_selected, d_steps, k_in.as<uint32_t>());
    ++ctx->launches;
    size_t tmp = 0;
    CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp, k_in.as<uint32_t>(), k_out.as<uint32_t>(), d_selected, d_ordered,
                                                  (int)n_selected, 0, 32, st));
    int rc = ensure_cub_temp(ctx, tmp);
    if (rc) return rc;
    CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(ctx->d_cub_temp, tmp, k_in.as<uint32_t>(), k_out.as<uint32_t>(), d_selected,
                                                  d_ordered, (int)n_selected, 0, 32, st));
    return GRRT_OK;
}

namespace {
// 8x4-pixel tiles dealt round-robin to the devices (tile-major, so 32 consecutive work units are one compact patch);
// the same partition as gr_raytracer_b200/renderer.py:tile_interleaved_items.  `order` (optional): the tiles in the
// order they are to be dealt and traced (a pilot pass puts the expensive ones first); such a partition is not cached.
int ensure_partition(GrrtCtx* ctx, uint32_t W, uint32_t H, const std::vector<uint32_t>* order = nullptr) {
    const size_t N = 1 + ctx->members.size();
    if (!order && ctx->part_w == W && ctx->part_h == H && ctx->d_part_items.size() == N) return GRRT_OK;
    for (size_t k = 0; k < ctx->d_part_items.size(); ++k) {
        if (!ctx->d_part_items[k]) continue;
        cudaSetDevice(k == 0 ? ctx->device : ctx->members[k - 1]->device);
        cudaFree(ctx->d_part_items[k]);
    }
    ctx->d_part_items.assign(N, nullptr);
    ctx->part_counts.assign(N, 0);
    ctx->part_w = ctx->part_h = 0;
    const uint64_t tiles_x = (W + 7) / 8, tiles_y = (H + 3) / 4, n_tiles = tiles_x * tiles_y;
    std::vector<uint32_t> ids;
    for (size_t d = 0; d < N; ++d) {
        ids.clear();
        for (uint64_t k = d; k < n_tiles; k += N) {
            const uint64_t t = order ? (*order)[k] : k;
            const uint64_t ty = t / tiles_x, tx = t % tiles_x;
            for (uint32_t q = 0; q < 32; ++q) {
                const uint64_t py = ty * 4 + (q >> 3), px = tx * 8 + (q & 7);
                if (py < H && px < W) ids.push_back((uint32_t)(py * W + px));
            }
        }
        GrrtCtx* c = d == 0 ? ctx : ctx->members[d - 1];
        CUDA_TRY(ctx, cudaSetDevice(c->device));
        ctx->part_counts[d] = ids.size();
        if (!ids.empty()) {
            CUDA_TRY(ctx, cudaMalloc(&ctx->d_part_items[d], ids.size() * 4));
            CUDA_TRY(ctx, cudaMemcpy(ctx->d_part_items[d], ids.data(), ids.size() * 4, cudaMemcpyHostToDevice));
        }
    }
    if (!order) {
        ctx->part_w = W;
        ctx->part_h = H;
    }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return GRRT_OK;
}

// ---- pilot pass (BASELINE.json north_star: "load-balanced from a pilot low-resolution pass") ---------------------------
// Frames with a large step budget have a heavy tail: in config 3 (Kerr-Schild, --max-steps 1e6) 8 % of the rays crawl
// towards the horizon for ~2.8e5 steps and 0.1 % use the whole budget, a serial chain of tens of seconds each.  Traced
// in image order those rays start in the middle of the pass and the pass ends long after the queue is empty.  The pilot
// traces the same section at 1/8 resolution with the step budget capped (a fraction of a per cent of the frame's work),
// and the 8x4 tiles are then traced -- and dealt to the devices -- in descending pilot cost, so the long rays start
// first.  The image does not depend on the order.
constexpr uint64_t kPilotMinBudget = 65536;     // scenes with a smaller --max-steps have no tail worth a pilot
constexpr uint64_t kPilotStepCap = 20000;
constexpr uint32_t kPilotScale = 8;
int pilot_tile_order(GrrtCtx* ctx, const GrrtDeviceScene* ds, uint32_t from_row, uint32_t from_col, uint32_t to_row, uint32_t to_col,
                     std::vector<uint32_t>* order) {
    const uint32_t W = to_col - from_col, H = to_row - from_row;
    GrrtDeviceScene pilot = *ds;   // same tables, coarser camera, capped budget
    DScene& p = pilot.s;
    const double rows = std::ceil(ds->s.cam_rows / kPilotScale), cols = std::ceil(ds->s.cam_cols / kPilotScale);
    p.cam_scale = ds->s.cam_scale * ds->s.cam_rows / rows;   // 2 tan(alpha/2) / rows (camera.rs:217-224)
    p.cam_rows = rows;
    p.cam_cols = cols;
    p.max_steps = std::min<uint64_t>(ds->s.max_steps, kPilotStepCap);
    const uint32_t fr = from_row / kPilotScale, fc = from_col / kPilotScale;
    const uint32_t tr = (to_row + kPilotScale - 1) / kPilotScale, tc = (to_col + kPilotScale - 1) / kPilotScale;
    const uint32_t PW = tc - fc, PH = tr - fr;
    const size_t pn = (size_t)PW * PH;
    cudaStream_t s = ctx->stream;
    DevBuf xyza, steps;
    CUDA_TRY(ctx, xyza.alloc(pn * 32, s));
    CUDA_TRY(ctx, steps.alloc(pn * 4, s));
    int rc = trace_items_impl(ctx, &pilot, nullptr, 0, pn, fr, fc, tr, tc, 0, xyza.as<double>(), nullptr, nullptr,
                              steps.as<uint32_t>(), nullptr /* not counted in the frame's statistics */, s, TraceOpts());
    if (rc) return rc;
    std::vector<uint32_t> cost(pn);
    CUDA_TRY(ctx, cudaMemcpyAsync(cost.data(), steps.p, pn * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaStreamSynchronize(s));
    const uint32_t tiles_x = (W + 7) / 8, tiles_y = (H + 3) / 4;
    std::vector<uint32_t> tile_cost((size_t)tiles_x * tiles_y);
    for (uint32_t ty = 0; ty < tiles_y; ++ty)
        for (uint32_t tx = 0; tx < tiles_x; ++tx) {
            // the tile's first pixel in frame coordinates -> the pilot pixel that covers it
            const uint32_t py = std::min((from_row + ty * 4) / kPilotScale - fr, PH - 1);
            const uint32_t px = std::min((from_col + tx * 8) / kPilotScale - fc, PW - 1);
            tile_cost[(size_t)ty * tiles_x + tx] = cost[(size_t)py * PW + px];
        }
    order->resize(tile_cost.size());
    for (size_t k = 0; k < order->size(); ++k) (*order)[k] = (uint32_t)k;
    std::stable_sort(order->begin(), order->end(), [&](uint32_t a, uint32_t b) { return tile_cost[a] > tile_cost[b]; });
    return GRRT_OK;
}

struct EventSet {   // events of one frame, destroyed on scope exit (each on the device it was created on)
    std::vector<std::pair<int, cudaEvent_t>> ev;
    ~EventSet() {
        for (auto& e : ev) {
            cudaSetDevice(e.first);
            cudaEventDestroy(e.second);
        }
    }
    cudaError_t make(int device, cudaEvent_t* out, unsigned flags = cudaEventDefault) {
        cudaError_t e = cudaSetDevice(device);
        if (e != cudaSuccess) return e;
        e = cudaEventCreateWithFlags(out, flags);
        if (e == cudaSuccess) ev.emplace_back(device, *out);
        return e;
    }
};
}  // namespace

// One frame: 1-spp base pass, adaptive mask, supersampling of the flagged pixels (raytracer.rs:177-193, 257-318), on the
// device of `ctx` or -- for a grrt_init_multi context -- sharded over all its devices.  Multi-device: every device
// traces its share of the 8x4 tiles and stores each finished pixel straight into the PRIMARY device's frame (peer
// stores over NVLink from inside the trace kernel; there is no reduction, hence no collective and no gather pass); the
// primary derives the mask and the cost order, every device supersamples every N-th flagged pixel of that order --
// reading the list from the primary's memory -- and stores the means into the primary's frame.  Stream-ordered events
// separate the passes; the host waits once (for the flagged count), as in the single-device frame.  The result is bit
// for bit the single-device frame.
int grrt_render_resident(GrrtCtx* ctx, const GrrtDeviceScene* ds, uint32_t from_row, uint32_t from_col, uint32_t to_row,
                         uint32_t to_col, double* xyza_out, uint8_t* ray_class_out, uint8_t* stop_reason_out,
                         uint32_t* steps_out, GrrtStats* stats) {
    if (!ctx || !ds || !xyza_out) return GRRT_ERR_INVALID_ARGUMENT;
    if (to_row < from_row || to_col < from_col) return fail(ctx, GRRT_ERR_INVALID_ARGUMENT, "inverted section");
    const size_t N = 1 + ctx->members.size();
    if (ds->replicas.size() != N - 1) return fail(ctx, GRRT_ERR_INVALID_ARGUMENT, "scene was not uploaded through this context");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc = GRRT_OK;
    const uint32_t W = to_col - from_col, H = to_row - from_row;
    const size_t n = (size_t)W * H;
    GrrtStats st;
    std::memset(&st, 0, sizeof st);
    if (n == 0) {
        if (stats) *stats = st;
        return GRRT_OK;
    }
    auto dev_ctx = [&](size_t d) { return d == 0 ? ctx : ctx->members[d - 1]; };
    auto dev_scene = [&](size_t d) { return d == 0 ? ds : ds->replicas[d - 1]; };
    struct Restore {   // leave the primary device current, whatever happens
        int device;
        ~Restore() { cudaSetDevice(device); }
    } restore{ctx->device};
    // a pilot pass decides the tile order of frames with a heavy tail (GRRT_PILOT=0 switches it off)
    const char* pilot_env = std::getenv("GRRT_PILOT");
    const bool use_pilot = ds->s.max_steps > kPilotMinBudget && n >= 4096 && !(pilot_env && pilot_env[0] == '0');
    EventSet events;
    cudaEvent_t ev_pilot = nullptr;
    if (use_pilot) {
        CUDA_TRY(ctx, events.make(ctx->device, &ev_pilot));
        CUDA_TRY(ctx, cudaEventRecord(ev_pilot, ctx->stream));
        std::vector<uint32_t> order;
        rc = pilot_tile_order(ctx, ds, from_row, from_col, to_row, to_col, &order);
        if (rc) return rc;
        rc = ensure_partition(ctx, W, H, &order);
        if (rc) return rc;
    } else if (N > 1) {
        rc = ensure_partition(ctx, W, H);
        if (rc) return rc;
    }
    cudaStream_t s = ctx->stream;
    DevBuf xyza, cls, stop, steps, sel, ordered, cnt;
    struct MembersIdle {   // destroyed before the buffers above: no member kernel may still store into a frame being freed
        GrrtCtx* ctx;
        ~MembersIdle() {
            for (GrrtCtx* m : ctx->members) {
                cudaSetDevice(m->device);
                cudaStreamSynchronize(m->stream);
            }
            cudaSetDevice(ctx->device);
        }
    } members_idle{ctx};
    CUDA_TRY(ctx, xyza.alloc(n * 32, s));
    CUDA_TRY(ctx, cls.alloc(n, s));
    CUDA_TRY(ctx, stop.alloc(n, s));
    CUDA_TRY(ctx, steps.alloc(n * 4, s));
    for (size_t d = 0; d < N; ++d) {
        GrrtCtx* c = dev_ctx(d);
        CUDA_TRY(ctx, cudaSetDevice(c->device));
        CUDA_TRY(ctx, cudaMemsetAsync(c->d_counters, 0, 16 * sizeof(unsigned long long), c->stream));
    }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaEvent_t ev[4];
    for (auto& e : ev) CUDA_TRY(ctx, events.make(ctx->device, &e));
    CUDA_TRY(ctx, cudaEventRecord(ev[0], s));
    // ---- base pass ----
    if (N == 1 && !use_pilot) {
        rc = grrt_trace_items_device(ctx, ds, nullptr, 0, n, from_row, from_col, to_row, to_col, 0, xyza.as<double>(),
                                     cls.as<uint8_t>(), stop.as<uint8_t>(), steps.as<uint32_t>(),
                                     (uint64_t*)ctx->d_counters, s);
        if (rc) return rc;
    } else {
        double* px[1] = {xyza.as<double>()};
        uint8_t* pc[1] = {cls.as<uint8_t>()};
        uint8_t* ps[1] = {stop.as<uint8_t>()};
        uint32_t* pt[1] = {steps.as<uint32_t>()};
        TraceOpts o;
        o.n_peers = 1;
        o.xyza_peers = px; o.class_peers = pc; o.stop_peers = ps; o.steps_peers = pt;
        for (size_t d = 0; d < N; ++d) {
            GrrtCtx* c = dev_ctx(d);
            if (d > 0) CUDA_TRY(ctx, cudaStreamWaitEvent(c->stream, ev[0], 0));   // the frame buffers exist from here on
            rc = trace_items_impl(c, dev_scene(d), ctx->d_part_items[d], 0, ctx->part_counts[d], from_row, from_col, to_row, to_col,
                                  0, px[0], nullptr, nullptr, nullptr, (uint64_t*)c->d_counters, c->stream, o);
            if (rc) { if (d > 0) ctx->last_error = c->last_error; return rc; }
            if (d > 0) {
                cudaEvent_t done;
                CUDA_TRY(ctx, events.make(c->device, &done, cudaEventDisableTiming));
                CUDA_TRY(ctx, cudaEventRecord(done, c->stream));
                CUDA_TRY(ctx, cudaSetDevice(ctx->device));
                CUDA_TRY(ctx, cudaStreamWaitEvent(s, done, 0));
            }
        }
        CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    }
    CUDA_TRY(ctx, cudaEventRecord(ev[1], s));
    st.rays_base = n;
    const GrrtAdaptive& ad = ds->adaptive;
    // Read-back overlapped with the second pass: when the caller's buffers are pinned, the base frame starts its way to the
    // host now, on a second stream, and only the pixels the second pass rewrites follow afterwards (packed).  Pageable
    // buffers would make the copy calls block this thread, so they take the plain path after the last pass.
    bool early_copy = false;
    struct CopyGuard {   // an early return must not leave copies into the caller's buffers in flight
        cudaStream_t st;
        const bool* started;
        ~CopyGuard() { if (*started) cudaStreamSynchronize(st); }
    } copy_guard{ctx->copy_stream, &early_copy};
    if (ad.enabled || ad.show_sampling_mask) {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, xyza_out) == cudaSuccess && pa.type == cudaMemoryTypeHost) early_copy = true;
        cudaGetLastError();
    }
    if (early_copy) {
        cudaStream_t cs = ctx->copy_stream;
        CUDA_TRY(ctx, cudaStreamWaitEvent(cs, ev[1], 0));
        CUDA_TRY(ctx, cudaMemcpyAsync(xyza_out, xyza.p, n * 32, cudaMemcpyDeviceToHost, cs));
        if (ray_class_out) CUDA_TRY(ctx, cudaMemcpyAsync(ray_class_out, cls.p, n, cudaMemcpyDeviceToHost, cs));
        if (stop_reason_out) CUDA_TRY(ctx, cudaMemcpyAsync(stop_reason_out, stop.p, n, cudaMemcpyDeviceToHost, cs));
        if (steps_out) CUDA_TRY(ctx, cudaMemcpyAsync(steps_out, steps.p, n * 4, cudaMemcpyDeviceToHost, cs));
    }
    uint32_t n_sel = 0;
    CUDA_TRY(ctx, cudaEventRecord(ev[2], s));
    if (ad.enabled || ad.show_sampling_mask) {  // raytracer.rs:184
        CUDA_TRY(ctx, sel.alloc(n * 4, s));
        CUDA_TRY(ctx, cnt.alloc(4, s));
        rc = grrt_select_mask_device(ctx, ds, xyza.as<double>(), cls.as<uint8_t>(), H, W, sel.as<uint32_t>(),
                                     cnt.as<uint32_t>(), &st.minimum_luminance, s);
        if (rc) return rc;
        CUDA_TRY(ctx, cudaMemcpyAsync(&n_sel, cnt.p, 4, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(ctx, cudaStreamSynchronize(s));
        st.pixels_supersampled = n_sel;
        if (n_sel) {
            if (ad.show_sampling_mask) {
                fill_mask_color_kernel<<<(n_sel + 255) / 256, 256, 0, s>>>(sel.as<uint32_t>(), n_sel, ad, xyza.as<double>());
                ++ctx->launches;
                CUDA_TRY(ctx, cudaGetLastError());
                CUDA_TRY(ctx, cudaEventRecord(ev[2], s));
            } else {
                if (ad.samples_per_axis == 0 || ad.samples_per_axis > 4096)
                    return fail(ctx, GRRT_ERR_INVALID_CONFIGURATION, "adaptive_sampling.samples_per_axis must be greater than zero");
                const uint32_t spa = (uint32_t)ad.samples_per_axis;
                CUDA_TRY(ctx, ordered.alloc((size_t)n_sel * 4, s));
                rc = grrt_order_by_cost_device(ctx, sel.as<uint32_t>(), n_sel, steps.as<uint32_t>(), ordered.as<uint32_t>(), s);
                if (rc) return rc;
                CUDA_TRY(ctx, cudaEventRecord(ev[2], s));
                if (N == 1) {
                    rc = grrt_trace_items_device(ctx, ds, ordered.as<uint32_t>(), 0, n_sel, from_row, from_col, to_row, to_col, spa,
                                                 xyza.as<double>(), nullptr, nullptr, nullptr, (uint64_t*)ctx->d_counters, s);
                    if (rc) return rc;
                } else {
                    double* px[1] = {xyza.as<double>()};
                    uint8_t* pc[1] = {cls.as<uint8_t>()};
                    uint8_t* ps[1] = {stop.as<uint8_t>()};
                    for (size_t d = 0; d < N && d < n_sel; ++d) {
                        GrrtCtx* c = dev_ctx(d);
                        TraceOpts o;
                        o.item_offset = d;
                        o.item_stride = N;
                        o.n_peers = 1;
                        o.xyza_peers = px; o.class_peers = pc; o.stop_peers = ps;
                        if (d > 0) CUDA_TRY(ctx, cudaStreamWaitEvent(c->stream, ev[2], 0));   // mask and order are final
                        const uint64_t mine = (n_sel - d + N - 1) / N;
                        rc = trace_items_impl(c, dev_scene(d), ordered.as<uint32_t>(), 0, mine, from_row, from_col, to_row, to_col,
                                              spa, px[0], nullptr, nullptr, nullptr, (uint64_t*)c->d_counters, c->stream, o);
                        if (rc) { if (d > 0) ctx->last_error = c->last_error; return rc; }
                        if (d > 0) {
                            cudaEvent_t done;
                            CUDA_TRY(ctx, events.make(c->device, &done, cudaEventDisableTiming));
                            CUDA_TRY(ctx, cudaEventRecord(done, c->stream));
                            CUDA_TRY(ctx, cudaSetDevice(ctx->device));
                            CUDA_TRY(ctx, cudaStreamWaitEvent(s, done, 0));
                        }
                    }
                    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
                }
                st.rays_supersample = (uint64_t)n_sel * spa * spa;
            }
        } else {
            CUDA_TRY(ctx, cudaEventRecord(ev[2], s));
        }
    }
    CUDA_TRY(ctx, cudaEventRecord(ev[3], s));
    DevBuf packed;
    uint32_t n_patch = 0;
    if (early_copy) {
        n_patch = n_sel;   // the second pass (supersampling or mask colour) touched exactly the selected pixels
        if (n_patch) {
            const size_t need = (size_t)n_patch * (4 + 32);
            if (need > ctx->h_stage_bytes) {
                if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
                ctx->h_stage = nullptr;
                ctx->h_stage_bytes = 0;
                CUDA_TRY(ctx, cudaMallocHost(&ctx->h_stage, need + need / 4));
                ctx->h_stage_bytes = need + need / 4;
            }
            CUDA_TRY(ctx, packed.alloc((size_t)n_patch * 32, s));
            gather_pixels_kernel<<<(n_patch + 255) / 256, 256, 0, s>>>(sel.as<uint32_t>(), n_patch, xyza.as<double>(), packed.as<double>());
            ++ctx->launches;
            CUDA_TRY(ctx, cudaGetLastError());
            CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_stage, sel.p, (size_t)n_patch * 4, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(ctx, cudaMemcpyAsync((char*)ctx->h_stage + (((size_t)n_patch * 4 + 7) & ~(size_t)7), packed.p, (size_t)n_patch * 32,
                                          cudaMemcpyDeviceToHost, s));
        }
    } else {
        CUDA_TRY(ctx, cudaMemcpyAsync(xyza_out, xyza.p, n * 32, cudaMemcpyDeviceToHost, s));
        if (ray_class_out) CUDA_TRY(ctx, cudaMemcpyAsync(ray_class_out, cls.p, n, cudaMemcpyDeviceToHost, s));
        if (stop_reason_out) CUDA_TRY(ctx, cudaMemcpyAsync(stop_reason_out, stop.p, n, cudaMemcpyDeviceToHost, s));
        if (steps_out) CUDA_TRY(ctx, cudaMemcpyAsync(steps_out, steps.p, n * 4, cudaMemcpyDeviceToHost, s));
    }
    unsigned long long c[16], total[16];
    std::memset(total, 0, sizeof total);
    for (size_t d = 0; d < N; ++d) {   // the primary's stream has waited for every member's kernels
        GrrtCtx* m = dev_ctx(d);
        CUDA_TRY(ctx, cudaSetDevice(m->device));
        CUDA_TRY(ctx, cudaMemcpyAsync(c, m->d_counters, sizeof c, cudaMemcpyDeviceToHost, m->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(m->stream));
        for (int k = 0; k < 16; ++k) total[k] += c[k];
    }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(s));
    if (early_copy) {
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->copy_stream));   // the base frame has landed; now the rewritten pixels on top
        const uint32_t* ids = static_cast<const uint32_t*>(ctx->h_stage);
        const double* vals = reinterpret_cast<const double*>((const char*)ctx->h_stage + (((size_t)n_patch * 4 + 7) & ~(size_t)7));
        for (uint32_t i = 0; i < n_patch; ++i) std::memcpy(xyza_out + 4ull * ids[i], vals + 4ull * i, 32);
    }
    st.attempts = total[CNT_ATTEMPTS];
    st.accepted_steps = total[CNT_ACCEPTED];
    st.raymarch_samples = total[CNT_SAMPLES];
    st.ray_errors = total[CNT_ERRORS];
    st.hit_stack_overflows = total[CNT_OVERFLOW];
    for (int k = 0; k < 6; ++k) st.stop_histogram[k] = total[CNT_STOP0 + k];
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev[0], ev[1]); st.device_ms_base = ms;
    cudaEventElapsedTime(&ms, ev[1], ev[2]); st.device_ms_mask = ms;
    cudaEventElapsedTime(&ms, ev[2], ev[3]); st.device_ms_supersample = ms;
    cudaEventElapsedTime(&ms, ev[0], ev[3]); st.device_ms_total = ms;
    if (use_pilot) {   // the pilot (its trace, the read-back of its step counts, the host-side tile sort) belongs to the base pass
        cudaEventElapsedTime(&ms, ev_pilot, ev[0]);
        st.device_ms_base += ms;
        st.device_ms_total += ms;
    }
    if (stats) *stats = st;
    return GRRT_OK;
}

static int trace_single(GrrtCtx* ctx, const GrrtScene* scene, const double* pos, const double* mom, int use_pixel,
                        int64_t row, int64_t col, GrrtStep* steps_out, uint64_t cap, uint64_t* n_out, int32_t* stop_reason) {
    if (!ctx || !scene || !n_out || !stop_reason) return GRRT_ERR_INVALID_ARGUMENT;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    GrrtDeviceScene* ds = nullptr;
    int rc = grrt_scene_upload(ctx, scene, &ds);
    if (rc) return rc;
    struct Guard {
        GrrtCtx* c;
        GrrtDeviceScene* d;
        ~Guard() { grrt_scene_free(c, d); }
    } guard{ctx, ds};
    if (!steps_out) cap = 0;
    DevBuf dsteps, meta;
    CUDA_TRY(ctx, dsteps.alloc(cap * sizeof(GrrtStep), ctx->stream));
    CUDA_TRY(ctx, meta.alloc(32, ctx->stream));
    unsigned long long* d_n = meta.as<unsigned long long>();
    int* d_stop = reinterpret_cast<int*>(d_n + 1);
    int* d_err = d_stop + 1;
    cudaStream_t s = ctx->stream;
    int lrc = grrt_exact::launch_trace_single(ds->s, pos, mom, use_pixel, (long long)row, (long long)col, dsteps.as<GrrtStep>(), cap,
                                              d_n, d_stop, d_err, s);
    if (lrc != 0) return fail(ctx, GRRT_ERR_CUDA, std::string("trace_single launch: ") + cudaGetErrorString((cudaError_t)lrc));
    ++ctx->launches;
    CUDA_TRY(ctx, cudaGetLastError());
    unsigned long long n = 0;
    int sr[2] = {0, 0};
    CUDA_TRY(ctx, cudaMemcpyAsync(&n, d_n, 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaMemcpyAsync(sr, d_stop, 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaStreamSynchronize(s));
    uint64_t m = std::min<uint64_t>(n, cap);
    if (m) {
        CUDA_TRY(ctx, cudaMemcpyAsync(steps_out, dsteps.p, m * sizeof(GrrtStep), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(ctx, cudaStreamSynchronize(s));
    }
    *n_out = n;
    *stop_reason = sr[0];
    return sr[1];  // GRRT_RAYERR_* (>0) of the ray, 0 when it integrated
}

int grrt_trace_ray(GrrtCtx* ctx, const GrrtScene* scene, const double pos[4], const double mom[4], GrrtStep* steps_out,
                   uint64_t cap, uint64_t* n_out, int32_t* stop_reason) {
    if (!pos || !mom) return GRRT_ERR_INVALID_ARGUMENT;
    return trace_single(ctx, scene, pos, mom, 0, 0, 0, steps_out, cap, n_out, stop_reason);
}
int grrt_trace_pixel_ray(GrrtCtx* ctx, const GrrtScene* scene, int64_t row, int64_t col, GrrtStep* steps_out,
                         uint64_t cap, uint64_t* n_out, int32_t* stop_reason) {
    return trace_single(ctx, scene, nullptr, nullptr, 1, row, col, steps_out, cap, n_out, stop_reason);
}

// test hook: the device's glibc-compatible pow on host arrays (tests/test_gpu_parity.py)
int grrt_debug_pow(GrrtCtx* ctx, const double* x, const double* y, double* out, size_t n) {
    if (!ctx || !x || !y || !out) return GRRT_ERR_INVALID_ARGUMENT;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    DevBuf dx, dy, dout;
    CUDA_TRY(ctx, dx.alloc(n * 8, ctx->stream));
    CUDA_TRY(ctx, dy.alloc(n * 8, ctx->stream));
    CUDA_TRY(ctx, dout.alloc(n * 8, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(dx.p, x, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(dy.p, y, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (grrt_exact::launch_debug_pow(dx.as<double>(), dy.as<double>(), dout.as<double>(), n, ctx->stream) != 0)
        return fail(ctx, GRRT_ERR_CUDA, "debug_pow launch failed");
    CUDA_TRY(ctx, cudaMemcpyAsync(out, dout.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return GRRT_OK;
}

// test hook: the kernels' sincos (csrc/grrt_device.cuh sincos_pi2) on host arrays
int grrt_debug_sincos(GrrtCtx* ctx, const double* x, double* sn, double* cs, size_t n) {
    if (!ctx || !x || !sn || !cs) return GRRT_ERR_INVALID_ARGUMENT;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    DevBuf dx, ds, dc;
    CUDA_TRY(ctx, dx.alloc(n * 8, ctx->stream));
    CUDA_TRY(ctx, ds.alloc(n * 8, ctx->stream));
    CUDA_TRY(ctx, dc.alloc(n * 8, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(dx.p, x, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (grrt_exact::launch_debug_sincos(dx.as<double>(), ds.as<double>(), dc.as<double>(), n, ctx->stream) != 0)
        return fail(ctx, GRRT_ERR_CUDA, "debug_sincos launch failed");
    CUDA_TRY(ctx, cudaMemcpyAsync(sn, ds.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(cs, dc.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return GRRT_OK;
}

int grrt_measure_fp64_peak(GrrtCtx* ctx, double* dfma_per_second, double* ms_out) {
    if (!ctx || !dfma_per_second) return GRRT_ERR_INVALID_ARGUMENT;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int blocks = ctx->num_sms * 8, threads = 256, iters = 20000;
    DevBuf out;
    CUDA_TRY(ctx, out.alloc((size_t)blocks * threads * 8, ctx->stream));
    cudaEvent_t e0, e1;
    CUDA_TRY(ctx, cudaEventCreate(&e0));
    CUDA_TRY(ctx, cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0, ctx->stream);
        dfma_chain_kernel<<<blocks, threads, 0, ctx->stream>>>(out.as<double>(), iters, 1.0000001, 1e-9);
        cudaEventRecord(e1, ctx->stream);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    CUDA_TRY(ctx, cudaGetLastError());
    double n_fma = (double)blocks * threads * (double)iters * 64.0;
    *dfma_per_second = n_fma / (best * 1e-3);
    if (ms_out) *ms_out = best;
    return GRRT_OK;
}

}  // extern "C"

namespace {
__global__ void cost_keys_kernel(const uint32_t* sel, uint32_t n, const uint32_t* steps, uint32_t* keys);
}
// forces the code of the small kernels onto the context's device (an attribute query loads a lazily loaded function)
static void preload_small_kernels(GrrtCtx* ctx) {
    cudaSetDevice(ctx->device);
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, luminance_keys_kernel);
    cudaFuncGetAttributes(&a, mask_flags_kernel);
    cudaFuncGetAttributes(&a, fill_mask_color_kernel);
    cudaFuncGetAttributes(&a, linear_max_kernel);
    cudaFuncGetAttributes(&a, tonemap_kernel);
    cudaFuncGetAttributes(&a, grrt_exact_resolve);
    cudaFuncGetAttributes(&a, cost_keys_kernel);
    cudaGetLastError();
}
