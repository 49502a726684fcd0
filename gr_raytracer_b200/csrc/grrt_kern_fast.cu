// grrt_kern_fast.cu -- the same kernels compiled with -fmad=true for the geometries whose right-hand side is
// analytic and smooth (Euclidean, EuclideanSpherical, Schwarzschild, KerrBL): contraction changes results at the
// 1e-16 level, the same size as the glibc-vs-CUDA sin/cos difference already present, and is verified against
// the oracle by the parity tests.  Kerr-Schild is not built here: it needs the exact kernels.
#define GRRT_NS grrt_fast
#define GRRT_FAST_MATH 1
#include "grrt_kern_impl.cuh"
