// Microbenchmark: does an FP64 warp instruction (16 lanes per SM sub-partition, 2 cycles in the pipe) leave the
// issue slot of its second cycle to other pipes?  Each thread runs 8 independent DFMA chains; between the DFMAs M
// independent integer / FP32 / shared-memory instructions are interleaved.  If time ~ max(2N, N + M) the slots are
// shared; if time ~ 2N + M the FP64 instruction blocks the scheduler for both cycles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_mix fp64_mix.cu && ./fp64_mix
#include <cstdio>
#include <cuda_runtime.h>

template <int M, int KIND>
__global__ void __launch_bounds__(256) mix(double* out, int iters, double a, double b, unsigned seed) {
    double x[8];
    unsigned u[8];
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-3 + i; u[i] = seed + threadIdx.x * 17 + i; f[i] = (float)i + threadIdx.x; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                x[i] = fma(x[i], a, b);
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const int j = (i + m) & 7;
                    if (KIND == 0) u[j] = (u[j] ^ (u[j] >> 3)) + 0x9e3779b9u;     // LOP3 + IADD
                    else if (KIND == 1) f[j] = fmaf(f[j], 1.0001f, 0.5f);          // FFMA
                    else u[j] = u[j] * 1664525u + 1013904223u;                    // IMAD
                }
            }
        }
    }
    double s = 0.0;
    unsigned t = 0;
    float g = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += x[i]; t += u[i]; g += f[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (double)t + (double)g;
}

template <int M, int KIND>
void run(const char* name, double* d, int sms) {
    const int iters = 2000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    mix<M, KIND><<<sms * 4, 256>>>(d, 10, 1.0000001, 1e-9, 1u);
    cudaEventRecord(e0);
    mix<M, KIND><<<sms * 4, 256>>>(d, iters, 1.0000001, 1e-9, 1u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    // per SM sub-partition: 4 CTAs x 8 warps / 4 = 8 warps; each executes iters*32 DFMA (+ M others per DFMA)
    const double dfma_warp_inst = 8.0 * iters * 32.0;
    const double cycles = ms * 1e-3 * 1.965e9;
    printf("%-6s M=%d: %.3f ms, cycles per DFMA per sub-partition %.3f (other instructions per DFMA: %d%s)\n", name, M, ms,
           cycles / dfma_warp_inst, KIND == 0 ? 2 * M : M, KIND == 0 ? " (LOP3+IADD)" : "");
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* d;
    cudaMalloc(&d, sizeof(double) * sms * 4 * 256);
    run<0, 0>("none", d, sms);
    run<1, 1>("ffma", d, sms);
    run<2, 1>("ffma", d, sms);
    run<3, 1>("ffma", d, sms);
    run<1, 2>("imad", d, sms);
    run<2, 2>("imad", d, sms);
    run<1, 0>("alu", d, sms);
    run<2, 0>("alu", d, sms);
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
