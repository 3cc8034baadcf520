// FFMA2 operand-form throughput: coefficient from a per-thread register vs a uniform (kernel-param) value.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float a, float b, const float* tab) {
    float2 x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f); y[i] = make_float2(i * 1e-3f, threadIdx.x * 1e-4f); }
    float ar = MODE == 0 ? tab[threadIdx.x & 31] : a;     // MODE 0: per-thread register; MODE 1: uniform
    float br = MODE == 0 ? tab[32 + (threadIdx.x & 31)] : b;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            // recurrence-like: x = a*x + y ; y = b*y + x  (two pair operands + one scalar)
            x[i] = __ffma2_rn(make_float2(ar, ar), x[i], y[i]);
            y[i] = __ffma2_rn(make_float2(br, br), y[i], x[i]);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y + y[i].x + y[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 1024 * 4);
    float* tab; cudaMalloc(&tab, 64 * 4);
    float h[64]; for (int i = 0; i < 64; ++i) h[i] = 0.999f; cudaMemcpy(tab, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) for (int warps = 4; warps <= 16; warps *= 2) {
        int iters = 20000;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, 0.999f, 0.998f, tab); else k<1><<<148, warps * 32>>>(out, iters, 0.999f, 0.998f, tab);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fma = 148.0 * warps * 32 * (double)iters * 32;
        printf("coef %s warps/SM %2d: %.3f ms  %.2f TFMA/s\n", mode ? "uniform " : "register", warps, ms, fma / ms / 1e9);
    }
    return 0;
}
