// cusumtools_b200 — shared device/host helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CT_OK 0
#define CT_ERR_ARG (-1)
#define CT_ERR_CUDA (-2)
#define CT_ERR_UNSUPPORTED (-3)
#define CT_ERR_CAPACITY (-4)

#define CT_WARP 32
#define CT_FULL 0xffffffffu

void ct_set_error(const char* fmt, ...);
int ct_check_launch(const char* what);
int ct_sm_count();
int ct_max_smem_optin();

// Launch counter: every kernel launch of this library goes through CT_COUNT_LAUNCH so
// bench.py can report `gpu_launches` (ct_launch_count / ct_launch_count_reset).
extern unsigned long long g_ct_launches;
#define CT_COUNT_LAUNCH() (++g_ct_launches)

static __device__ __forceinline__ int ct_lane() { return threadIdx.x & 31; }

// 128-bit streaming loads/stores (read-once / write-once data: keep it out of L1).
static __device__ __forceinline__ uint4 ct_ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
static __device__ __forceinline__ void ct_stg_stream(void* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// -DCT_BOUNDS_CHECK: every hand-computed offset into a workspace / table is checked against the extent the host
// passed along (compute-sanitizer is not available on every pool); a violation prints the site and traps.  The
// production build compiles the checks away.
#ifdef CT_BOUNDS_CHECK
#define CT_CHECK_RANGE(off, count, limit, what)                                                                     \
    do {                                                                                                            \
        const long long ct_o_ = (long long)(off), ct_l_ = (long long)(limit);                                       \
        if (ct_o_ < 0 || ct_o_ + (long long)(count) > ct_l_) {                                                      \
            printf("CT_BOUNDS_CHECK %s: offset %lld + %d beyond %lld (block %d thread %d)\n", what, ct_o_, (int)(count), ct_l_, \
                   (int)blockIdx.x, (int)threadIdx.x);                                                              \
            asm volatile("trap;");                                                                                  \
        }                                                                                                           \
    } while (0)
#else
#define CT_CHECK_RANGE(off, count, limit, what) do { } while (0)
#endif
