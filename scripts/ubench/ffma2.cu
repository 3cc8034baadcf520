// Microbenchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) issue rate on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
    float2 x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    float2 aa = make_float2(a, a), bb = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { x[i].x = fmaf(x[i].x, a, b); x[i].y = fmaf(x[i].y, a, b); }
            else x[i] = __ffma2_rn(x[i], aa, bb);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) for (int warps = 4; warps <= 32; warps *= 2) {
        int iters = 20000;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, 0.999f, 0.001f); else k<1><<<148, warps * 32>>>(out, iters, 0.999f, 0.001f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fma = 148.0 * warps * 32 * (double)iters * 16;
        printf("mode %s warps/SM %2d: %.3f ms  %.2f TFMA/s  (%.1f FMA/clk/SM at 1.9 GHz)\n", mode ? "FFMA2" : "FFMA ", warps, ms, fma / ms / 1e9, fma / ms / 1e9 * 1e12 / 148 / 1.9e9 / 1e0 / 1e0 * 1e-0 / 1e0);
    }
    return 0;
}
