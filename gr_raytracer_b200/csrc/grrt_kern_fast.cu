// grrt_kern_fast.cu -- the same kernels compiled with -fmad=true for the geometries whose right-hand side is
// analytic and smooth (Euclidean, EuclideanSpherical, Schwarzschild, KerrBL): contraction changes results at the
// 1e-16 level, the same size as the glibc-vs-CUDA sin/cos difference already present, and is verified against
// the oracle by the parity tests.  Kerr-Schild is not built here: it needs the exact kernels.
#define GRRT_NS grrt_fast
#define GRRT_FAST_MATH 1
// the six RKF45 stages as straight-line code (their right-hand sides are ~100 instructions each here; the exact build
// keeps the stage loop because of the 2000-instruction Kerr-Schild right-hand side), and sin/cos of a stage's polar
// angle from the state's by the addition theorem (grrt_trace.cuh)
#ifndef GRRT_NO_UNROLL_STAGES
#define GRRT_UNROLL_STAGES 1
#endif
#ifndef GRRT_NO_SINCOS_DELTA
#define GRRT_SINCOS_DELTA 1
#endif
// y_new and the error vector of the stored components are formed at the last stage from k1..k5 in shared memory instead of
// being carried in registers through all six stages (C2 -2.4 %, C1 -1.6 %; no gain for the exact Kerr-Schild kernel)
#ifndef GRRT_NO_DEFER_SUMS
#define GRRT_DEFER_SUMS 1
#endif
#include "grrt_kern_impl.cuh"
