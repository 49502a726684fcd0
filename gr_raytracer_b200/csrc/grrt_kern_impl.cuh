// grrt_kern_impl.cuh -- instantiates the kernels of grrt_trace.cuh inside namespace GRRT_NS.
// Included by grrt_kern_exact.cu (compiled with -fmad=false: the reference's rounding, operation by
// operation) and grrt_kern_fast.cu (compiled with -fmad=true: a*b+c contracts to one DFMA).
#pragma once

#include "grrt_trace.cuh"

#ifndef GRRT_MB
#define GRRT_MB 4  // resident CTAs per SM the register allocator is asked to allow for the non-volumetric kernels
#endif

namespace GRRT_NS {

#ifndef GRRT_MB_BL
#define GRRT_MB_BL 5  // KerrBL: 96 registers, 5 CTAs/SM measured faster than 4 (and than 6); Schwarzschild prefers 4
#endif
#ifndef GRRT_MBV
#define GRRT_MBV 3  // resident CTAs/SM the volumetric kernels are register-allocated for
#endif
#ifndef GRRT_MB_KS
#define GRRT_MB_KS 4   // Kerr-Schild: resident CTAs/SM the register allocator targets (128 registers; 3 CTAs / 168 registers
                       // measured 8 % slower on a 640x640 frame, 5 and 6 CTAs slower still)
#endif
#ifndef GRRT_MBV_KS
#define GRRT_MBV_KS 2  // ... with a VolumetricDisc in the scene
#endif
template <bool V>
int launch_trace_v(const DScene& s, const TraceArgs& a, const TraceLaunch& l) {
    switch (s.geom) {
        case GRRT_GEOM_EUCLIDEAN: return launch_trace_t<GRRT_GEOM_EUCLIDEAN, V, (V ? GRRT_MBV : 4)>(s, a, l);
        case GRRT_GEOM_EUCLIDEAN_SPHERICAL: return launch_trace_t<GRRT_GEOM_EUCLIDEAN_SPHERICAL, V, (V ? GRRT_MBV : 4)>(s, a, l);
        case GRRT_GEOM_SCHWARZSCHILD: return launch_trace_t<GRRT_GEOM_SCHWARZSCHILD, V, (V ? GRRT_MBV : GRRT_MB)>(s, a, l);
        case GRRT_GEOM_KERR_BL: return launch_trace_t<GRRT_GEOM_KERR_BL, V, (V ? GRRT_MBV : GRRT_MB_BL)>(s, a, l);
#ifdef GRRT_WITH_KERR_SCHILD
        case GRRT_GEOM_KERR: return launch_trace_t<GRRT_GEOM_KERR, V, (V ? GRRT_MBV_KS : GRRT_MB_KS)>(s, a, l);
#endif
        default: return (int)cudaErrorInvalidValue;
    }
}

int launch_trace(const DScene& s, const TraceArgs& a, bool has_vol, const TraceLaunch& l) {
    return has_vol ? launch_trace_v<true>(s, a, l) : launch_trace_v<false>(s, a, l);
}

#ifdef GRRT_EXACT_EXTRAS
// ---- single-ray trajectory export (integrator.rs:78-174 with every step stored) ---------------------
template <int G>
__global__ void trace_single_kernel(const __grid_constant__ DScene S, const double4 pos4, const double4 mom4,
                                    int use_pixel, long long row, long long col, GrrtStep* steps,
                                    unsigned long long cap, unsigned long long* n_out, int* stop_out, int* err_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double pos[4] = {pos4.x, pos4.y, pos4.z, pos4.w}, mom[4] = {mom4.x, mom4.y, mom4.z, mom4.w};
    if (use_pixel) camera_ray<G>(S, (double)row, (double)col, pos, mom);
    double y[8];
    RayConst C;
    Freq F;
    init_ray<G>(S, pos, mom, y, C, F);
    unsigned long long n = 0;
    double t = 0.0;
    auto store = [&](unsigned long long step) {
        if (n < cap) {
            double p[4];
            momentum_from_state<G>(S, C, y, p);
            for (int i = 0; i < 4; ++i) {
                steps[n].x[i] = y[i];
                steps[n].p[i] = p[i];
            }
            steps[n].t = t;
            steps[n].step = step;
        }
        ++n;
    };
    store(0);
    double h = S.step_size;
    int stop = GRRT_STOP_NONE, error = 0;
    const PowTables pt = pow_tables_global();
    for (unsigned long long i = 1; i < S.max_steps && !error; ++i) {
        double h_cur = rclamp(h, kHMin, kHMax);
        bool ok = false;
        double y_new[8], err;
        for (int r = 0; r < kMaxRetryStep; ++r) {
            rkf45_step<G>(S, C, y, h_cur, y_new, err);
            double h_proposed = err > 0.0 ? 0.9 * h_cur * pow_glibc(pt, S.epsilon / err, 1.0 / 5.0) : h_cur * kHGrowthCap;
            h_proposed = rclamp(fmin(h_proposed, h_cur * kHGrowthCap), kHMin, kHMax);
            if (err > S.epsilon) {
                if (h_cur <= kHMin) {
                    h = h_cur;
                    ok = true;
                    break;
                }
                h_cur = rclamp(h_proposed / 2.0, kHMin, kHMax);
            } else {
                h = (err / S.epsilon < 1e-5) ? rclamp(h_cur * kHGrowthCap, kHMin, kHMax) : h_proposed;
                ok = true;
                break;
            }
        }
        if (!ok) {
            error = GRRT_RAYERR_RKF45_RETRIES;
            stop = GRRT_STOP_ERROR;
            break;
        }
        for (int k = 0; k < 8; ++k) y[k] = y_new[k];
        t += h_cur;
        store(i);
        stop = should_stop<G>(S, y, position_cartesian<G>(S, y), i);
        if (stop != GRRT_STOP_NONE) break;
    }
    *n_out = n;
    *stop_out = stop;
    *err_out = error;
}

int launch_trace_single(const DScene& s, const double pos[4], const double mom[4], int use_pixel, long long row, long long col,
                        GrrtStep* d_steps, unsigned long long cap, unsigned long long* d_n, int* d_stop, int* d_err,
                        cudaStream_t stream) {
    double4 p4 = make_double4(0, 0, 0, 0), m4 = make_double4(0, 0, 0, 0);
    if (!use_pixel) {
        p4 = make_double4(pos[0], pos[1], pos[2], pos[3]);
        m4 = make_double4(mom[0], mom[1], mom[2], mom[3]);
    }
#define LAUNCH_SINGLE(G) trace_single_kernel<G><<<1, 32, 0, stream>>>(s, p4, m4, use_pixel, row, col, d_steps, cap, d_n, d_stop, d_err)
    switch (s.geom) {
        case GRRT_GEOM_EUCLIDEAN: LAUNCH_SINGLE(GRRT_GEOM_EUCLIDEAN); break;
        case GRRT_GEOM_EUCLIDEAN_SPHERICAL: LAUNCH_SINGLE(GRRT_GEOM_EUCLIDEAN_SPHERICAL); break;
        case GRRT_GEOM_SCHWARZSCHILD: LAUNCH_SINGLE(GRRT_GEOM_SCHWARZSCHILD); break;
        case GRRT_GEOM_KERR: LAUNCH_SINGLE(GRRT_GEOM_KERR); break;
        case GRRT_GEOM_KERR_BL: LAUNCH_SINGLE(GRRT_GEOM_KERR_BL); break;
        default: return (int)cudaErrorInvalidValue;
    }
#undef LAUNCH_SINGLE
    return (int)cudaGetLastError();
}

__global__ void debug_pow_kernel(const double* x, const double* y, double* out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = pow_glibc(pow_tables_global(), x[i], y[i]);
}
int launch_debug_pow(const double* x, const double* y, double* out, size_t n, cudaStream_t stream) {
    if (n) debug_pow_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(x, y, out, n);
    return (int)cudaGetLastError();
}
__global__ void debug_sincos_kernel(const double* x, double* sn, double* cs, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) sincos_pi2(x[i], sn + i, cs + i);
}
int launch_debug_sincos(const double* x, double* sn, double* cs, size_t n, cudaStream_t stream) {
    if (n) debug_sincos_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(x, sn, cs, n);
    return (int)cudaGetLastError();
}
#endif  // GRRT_EXACT_EXTRAS

}  // namespace GRRT_NS
