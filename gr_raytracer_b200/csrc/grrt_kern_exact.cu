// grrt_kern_exact.cu -- kernels compiled with -fmad=false: every + - * / sqrt rounds exactly as in the
// reference's Rust build (no FMA contraction).  All five geometries; mandatory for Kerr-Schild, whose
// finite-difference right-hand side amplifies last-bit differences (see grrt_device.cuh).
#define GRRT_NS grrt_exact
#define GRRT_WITH_KERR_SCHILD 1
#define GRRT_EXACT_EXTRAS 1
#include "grrt_kern_impl.cuh"
