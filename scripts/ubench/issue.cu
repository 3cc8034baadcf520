// Does an FFMA2 (rt = 2 on the fma pipe) block the SMSP's issue port for its second cycle?
// Interleaves independent FFMA2 chains with independent ALU-pipe ops (LOP3/IADD3) at ratios 1:1 and 1:2.
#include <cstdio>
#include <cuda_runtime.h>
template <int NALU, int NF2, int NF1>
__global__ void k(float* out, int iters, float a, int m, long long* cyc) {
    float2 x[8]; int q[8]; float s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f); q[i] = threadIdx.x * 7 + i; s[i] = i + threadIdx.x; }
    const float2 c = make_float2(a, a), d = make_float2(0.001f, 0.002f);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (NF2 > 0) x[i] = __ffma2_rn(c, x[i], d);
            if (NF2 > 1) x[i] = __ffma2_rn(d, x[i], c);
            if (NF1 > 0) s[i] = fmaf(s[i], x[(i + 1) & 7].x, s[(i + 3) & 7]);
            if (NALU > 0) q[i] = (q[i] ^ m) + it;
            if (NALU > 1) q[i] = (q[i] & m) - it;
        }
    }
    long long t1 = clock64();
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += x[i].x + x[i].y + q[i] + s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NALU, int NF2, int NF1> void run(float* out, long long* cyc) {
    const int warps = 16, iters = 4000;
    for (int rep = 0; rep < 2; ++rep) { k<NALU, NF2, NF1><<<148, warps * 32>>>(out, iters, 0.999f, 0x5555, cyc); cudaDeviceSynchronize(); }
    double groups = (double)iters * 8 * (warps / 4);
    printf("per group: %d FFMA2 + %d FFMA(3-reg) + %d ALU-pairs : %.2f cycles/group/SMSP\n", NF2, NF1, NALU, (double)*cyc / groups);
}
int main() {
    float* out; cudaMalloc(&out, 148 * 1024 * 4);
    long long* cyc; cudaMallocManaged(&cyc, 8);
    run<0, 1, 0>(out, cyc); run<1, 0, 0>(out, cyc); run<2, 0, 0>(out, cyc); run<1, 1, 0>(out, cyc); run<2, 1, 0>(out, cyc); run<2, 2, 0>(out, cyc);
    run<0, 0, 1>(out, cyc); run<0, 1, 1>(out, cyc); run<1, 0, 1>(out, cyc);
    return 0;
}
