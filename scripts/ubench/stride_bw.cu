// DRAM bandwidth of the filter passes' access pattern: many concurrent streams, each a contiguous region of L bytes,
// touched G bytes at a time ("lane = run": a warp owns 64 adjacent regions and moves G bytes of each per iteration).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o stride_bw stride_bw.cu && ./stride_bw
// Prints GB/s for reads, writes and a 1:4 read:write mix at G = 64 .. 2048 and for the plain streaming order.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int MODE, int G16>   // 0 read, 1 write, 2 read 1 part (a second buffer, same pattern at quarter size) + write 4 parts
__global__ void __launch_bounds__(128, 3) k_stride(const uint4* __restrict__ in, uint4* __restrict__ out, long long ngroups, int L16,
                                                   unsigned long long* next, uint4* sink) {
    // region = L16 uint4; a warp's group = 64 regions; per iteration G16 uint4 of each region
    const int lane = threadIdx.x & 31;
    uint4 acc = make_uint4(0, 0, 0, 0);
    constexpr int per_row = G16;                   // uint4 per row per iteration
    for (;;) {
        unsigned long long g = 0;
        if (lane == 0) g = atomicAdd(next, 1ULL);
        g = __shfl_sync(0xffffffffu, g, 0);
        if ((long long)g >= ngroups) break;
        const long long base = (long long)g * 64 * L16;
        for (int t = L16 / G16 - 1; t >= 0; --t) {
            // 64 rows x G16 uint4: lanes walk (row, col) pairs, consecutive lanes on consecutive 16-byte units of a row
            constexpr int NIT = 64 * G16 / 32;
            uint4 v[NIT < 16 ? NIT : 16];
#pragma unroll 1
            for (int i0 = 0; i0 < NIT; i0 += 16) {
#pragma unroll
                for (int u = 0; u < 16 && u < NIT; ++u) {
                    const int i = (i0 + u) * 32 + lane;
                    const int row = i / per_row, col = i % per_row;
                    const long long off = base + (long long)row * L16 + (long long)t * G16 + col;
                    if (MODE == 0) v[u] = __ldcs(in + off);
                    else if (MODE == 1) __stcs(out + off, make_uint4(i, t, lane, (unsigned)g));
                    else {
                        if ((col & 3) == 0) v[u] = __ldcs(in + (off >> 2)); else v[u] = make_uint4(0, 0, 0, 0);
                        __stcs(out + off, make_uint4(i, t, lane, (unsigned)g));
                    }
                }
                if (MODE != 1) {
#pragma unroll
                    for (int u = 0; u < 16 && u < NIT; ++u) { acc.x ^= v[u].x; acc.y += v[u].y; acc.z ^= v[u].z; acc.w += v[u].w; }
                }
            }
        }
    }
    if (acc.x == 0x12345678u && acc.y == 77u) *sink = acc;
}

int main() {
    const long long bytes = 4LL << 30;
    uint4 *a, *b, *sink; unsigned long long* ctr;
    cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&sink, 64); cudaMalloc(&ctr, 8);
    cudaMemset(a, 1, bytes); cudaMemset(b, 2, bytes);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int Ls[] = {8192, 16384};
    for (int li = 0; li < 2; ++li) {
        const int L = Ls[li];
        for (int G = 64; G <= L; G *= 2) {
            if (G > 2048 && G != L) continue;
            const int L16 = L / 16, G16 = G / 16;
            const long long ngroups = bytes / (64LL * L);
            float ms[3];
            for (int mode = 0; mode < 3; ++mode) {
                float best = 1e9f;
                for (int rep = 0; rep < 3; ++rep) {
                    cudaMemset(ctr, 0, 8);
                    cudaEventRecord(e0);
#define LAUNCH(GG) { if (mode == 0) k_stride<0, GG><<<sms * 3, 128>>>(a, b, ngroups, L16, ctr, sink); \
                     if (mode == 1) k_stride<1, GG><<<sms * 3, 128>>>(a, b, ngroups, L16, ctr, sink); \
                     if (mode == 2) k_stride<2, GG><<<sms * 3, 128>>>(a, b, ngroups, L16, ctr, sink); }
                    switch (G16) { case 4: LAUNCH(4) break; case 8: LAUNCH(8) break; case 16: LAUNCH(16) break; case 32: LAUNCH(32) break;
                                   case 64: LAUNCH(64) break; case 128: LAUNCH(128) break; case 512: LAUNCH(512) break; case 1024: LAUNCH(1024) break; }
                    cudaEventRecord(e1); cudaEventSynchronize(e1);
                    float t; cudaEventElapsedTime(&t, e0, e1); if (t < best) best = t;
                }
                ms[mode] = best;
            }
            const double gb = bytes / 1e9;
            printf("region %5d B, %4d B at a time (%d streams in flight): read %6.0f GB/s  write %6.0f GB/s  read 1 : write 4 %6.0f GB/s\n",
                   L, G, sms * 12 * 64, gb / ms[0] * 1e3, gb / ms[1] * 1e3, gb * 1.25 / ms[2] * 1e3);
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
