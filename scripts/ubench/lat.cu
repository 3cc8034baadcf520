// Dependent-chain latency and per-warp issue rate of FFMA vs FFMA2 on sm_100a.
// ILP independent chains per thread, W warps per SM (W/4 per SMSP); prints cycles per
// instruction per warp and the implied latency (cycles per round of ILP instructions).
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, bool PAIR>
__global__ void k(float* out, int iters, float a, float b, long long* cyc) {
    float2 x[ILP]; float s[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f); s[i] = threadIdx.x + i; }
    const float2 c = make_float2(a, a), d = make_float2(b, b);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (PAIR) x[i] = __ffma2_rn(c, x[i], d); else s[i] = fmaf(a, s[i], b);
            }
    }
    long long t1 = clock64();
    float r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r += x[i].x + x[i].y + s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP, bool PAIR> void run(float* out, long long* cyc) {
    for (int warps = 4; warps <= 16; warps += 4) {
        const int iters = 2000;
        for (int rep = 0; rep < 2; ++rep) { k<ILP, PAIR><<<148, warps * 32>>>(out, iters, 0.999f, 0.001f, cyc); cudaDeviceSynchronize(); }
        double rounds = (double)iters * 8;
        printf("%s ILP %2d warps/SMSP %d: %.2f cycles per round (=latency if < pipe), %.3f warp-instr/clk/SMSP\n", PAIR ? "FFMA2" : "FFMA ",
               ILP, warps / 4, (double)*cyc / rounds, rounds * ILP * (warps / 4) / (double)*cyc);
    }
}
int main() {
    float* out; cudaMalloc(&out, 148 * 1024 * 4);
    long long* cyc; cudaMallocManaged(&cyc, 8);
    run<1, false>(out, cyc); run<2, false>(out, cyc); run<4, false>(out, cyc); run<8, false>(out, cyc);
    run<1, true>(out, cyc); run<2, true>(out, cyc); run<4, true>(out, cyc); run<8, true>(out, cyc);
    return 0;
}
