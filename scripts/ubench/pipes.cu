// Issue-rate microbenchmarks that size the filter / CUSUM kernels on sm_100a:
//   SHFL alone, FFMA2 alone, FFMA2 + SHFL interleaved, FFMA2 + FFMA interleaved, LDS.32 alone,
//   I2F / F2I conversions.  Reports warp-instructions per clock per SM (clock64-based).
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, float a, long long* cyc) {
    __shared__ float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 1e-3f;
    __syncthreads();
    float2 x[8];
    float s[8];
    int q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f); s[i] = threadIdx.x + i; q[i] = threadIdx.x * 7 + i; }
    const float2 c = make_float2(a, a);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) s[i] = __shfl_up_sync(0xffffffffu, s[i], 1);
            if (MODE == 1) x[i] = __ffma2_rn(c, x[i], x[i]);
            if (MODE == 2) { x[i] = __ffma2_rn(c, x[i], x[i]); s[i] = __shfl_up_sync(0xffffffffu, s[i], 1); }
            if (MODE == 3) { x[i] = __ffma2_rn(c, x[i], x[i]); s[i] = fmaf(s[i], a, s[i]); }
            if (MODE == 4) s[i] += sm[(threadIdx.x + i * 32 + it) & 4095];
            if (MODE == 5) { s[i] = (float)q[i]; q[i] = q[i] + it; }
            if (MODE == 6) { q[i] = __float2int_rn(s[i]); s[i] = s[i] + a; }
            if (MODE == 7) s[i] = fmaf(s[i], a, s[i]);
            if (MODE == 8) { x[i] = __ffma2_rn(c, x[i], x[i]); s[i] += sm[(threadIdx.x + i * 32 + it) & 4095]; }
        }
    }
    long long t1 = clock64();
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += x[i].x + x[i].y + s[i] + q[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    float* out; cudaMalloc(&out, 148 * 1024 * 4);
    long long* cyc; cudaMallocManaged(&cyc, 8);
    const char* names[] = {"SHFL", "FFMA2", "FFMA2+SHFL", "FFMA2+FFMA", "LDS.32+FADD", "I2F+IADD", "F2I+FADD", "FFMA", "FFMA2+LDS+FADD"};
    const int per[] = {1, 1, 2, 2, 2, 2, 2, 1, 3};
    for (int mode = 0; mode < 9; ++mode)
        for (int warps = 4; warps <= 16; warps *= 2) {
            const int iters = 4000;
            for (int rep = 0; rep < 2; ++rep) {
                switch (mode) {
                    case 0: k<0><<<148, warps * 32>>>(out, iters, 0.999f, cyc); break;
                    case 1: k<1><<<148, warps * 32>>>(out, iters, 0.999f, cyc); break;
                    case 2: k<2><<<148, warps * 32>>>(out, iters, 0.999f, cyc); break;
                    case 3: k<3><<<148, warps * 32>>>(out, iters, 0.999f, cyc); break;
                    case 4: k<4><<<148, warps * 32>>>(out, iters, 0.999f, cyc); break;
                    case 5: k<5><<<148, warps * 32>>>(out, iters, 0.999f, cyc); break;
                    case 6: k<6><<<148, warps * 32>>>(out, iters, 0.999f, cyc); break;
                    case 7: k<7><<<148, warps * 32>>>(out, iters, 0.999f, cyc); break;
                    case 8: k<8><<<148, warps * 32>>>(out, iters, 0.999f, cyc); break;
                }
                cudaDeviceSynchronize();
            }
            const double winstr = (double)warps * iters * 8 * per[mode];
            printf("%-16s warps/SM %2d: %9lld cycles  %.3f warp-instr/clk/SM (listed ops only)\n", names[mode], warps, *cyc,
                   winstr / (double)*cyc);
        }
    return 0;
}
