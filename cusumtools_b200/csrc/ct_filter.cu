// C-ABI entry points of stage 1 (dequantise + median pad + zero-phase Bessel; plot-trace.py:272-287, 313-320)
// and the exact median of the masked codes that np.pad(mode='median') needs (plot-trace.py:319).
// The filter passes themselves live in ct_filter_seq.cu.
#include "ct_common.cuh"
#include "cusumtools_b200.h"

namespace {
// ------------------------------ exact median of u16 codes ----------------------------
// Sampled histogram (estimate) + exact window count (verification); the host replays the
// reference's scale_raw_data on the selected code(s) so the pad value is bit-identical
// to np.median(scale_raw_data(raw)).
__global__ void ct_hist_sampled_kernel(const uint16_t* __restrict__ raw, long long n, long long stride,
                                       unsigned mask, unsigned* __restrict__ hist) {
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * stride;
    const long long step = (long long)gridDim.x * blockDim.x * stride;
    for (; i < n; i += step) atomicAdd(&hist[raw[i] & mask], 1u);
}

constexpr int kWin = 8;
// out[0] = #codes < lo ; out[1+i] = #codes == lo + i*step (step = power of two, lo a multiple of it).
// Eight 8-bit in-register counters in two 32-bit words: with d = code - lo (negative below the
// window) the increment is 1 << (8 * d/step) through PTX shl.b32, which CLAMPS shift amounts >= 32 to
// "all bits out" (result 0), so codes below or above the window add nothing without any compare; the
// sign bit of d is the "below" count.  ~10 integer operations per code.
static __device__ __forceinline__ unsigned shl_clamp(unsigned v, unsigned amt) {
    unsigned r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(amt));
    return r;
}
// WIDE = false: only the first four window codes are counted (one counter word: 7 instead of 10 operations per code).
template <bool WIDE>
__global__ void __launch_bounds__(256)
ct_count_window_kernel(const uint16_t* __restrict__ raw, long long n, unsigned mask, unsigned lo,
                       unsigned step, unsigned long long* __restrict__ out /*[1+kWin]*/) {
    const int sh = __ffs(step) - 1;
    unsigned long long tot[1 + kWin];
#pragma unroll
    for (int i = 0; i <= kWin; ++i) tot[i] = 0;
    unsigned below = 0, c0 = 0, c1 = 0;
    auto tally = [&](unsigned code) {                      // code already masked
        const unsigned d = code - lo;
        below += d >> 31;                                  // codes and lo are < 2^16: negative iff code < lo
        const unsigned amt = sh >= 3 ? d >> (sh - 3) : d << (3 - sh);   // 8 * (d / step): d is a multiple of step
        c0 += shl_clamp(1u, amt);
        if (WIDE) c1 += shl_clamp(1u, amt - 32u);
    };
    auto flush = [&]() {
        tot[0] += below; below = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) { tot[1 + i] += (c0 >> (8 * i)) & 0xffu; tot[5 + i] += (c1 >> (8 * i)) & 0xffu; }
        c0 = 0; c1 = 0;
    };
    const unsigned m2 = mask | (mask << 16);
    const long long nvec = n / 8;
    const uint4* v = reinterpret_cast<const uint4*>(raw);
    const bool aligned = (reinterpret_cast<uintptr_t>(raw) & 15) == 0;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nth = (long long)gridDim.x * blockDim.x;
    if (aligned) {
        long long i = tid;
        int rounds = 0;
        // four independent 16-byte loads per batch, the next batch on its way while this one is tallied
        uint4 q[4], nq[4];
        bool have = i + 3 * nth < nvec;
        if (have) {
#pragma unroll
            for (int u = 0; u < 4; ++u) q[u] = ct_ldg_stream(v + i + u * nth);
        }
        while (have) {
            const long long inext = i + 4 * nth;
            const bool more = inext + 3 * nth < nvec;
            if (more) {
#pragma unroll
                for (int u = 0; u < 4; ++u) nq[u] = ct_ldg_stream(v + inext + u * nth);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned ww[4] = {q[u].x & m2, q[u].y & m2, q[u].z & m2, q[u].w & m2};
#pragma unroll
                for (int j = 0; j < 4; ++j) { tally(ww[j] & 0xffffu); tally(ww[j] >> 16); }
            }
            if (++rounds == 7) { flush(); rounds = 0; }     // 32 tallies per round: an 8-bit counter holds 7 rounds
#pragma unroll
            for (int u = 0; u < 4; ++u) q[u] = nq[u];
            i = inext;
            have = more;
        }
        flush();
        for (; i < nvec; i += nth) {
            const uint4 q = ct_ldg_stream(v + i);
            const unsigned ww[4] = {q.x & m2, q.y & m2, q.z & m2, q.w & m2};
#pragma unroll
            for (int j = 0; j < 4; ++j) { tally(ww[j] & 0xffffu); tally(ww[j] >> 16); }
            flush();
        }
        for (long long k = nvec * 8 + tid; k < n; k += nth) { tally(raw[k] & mask); flush(); }
    } else {
        int since = 0;
        for (long long k = tid; k < n; k += nth) { tally(raw[k] & mask); if (++since == 128) { flush(); since = 0; } }
    }
    flush();
#pragma unroll
    for (int i = 0; i < 1 + (WIDE ? kWin : kWin / 2); ++i) {
        unsigned long long sum = tot[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(CT_FULL, sum, o);
        if (ct_lane() == 0 && sum) atomicAdd(&out[i], sum);
    }
}


}  // namespace

// lane-sequential two-pass implementation (ct_filter_seq.cu)
int ct_filtfilt_seq(const void* in, int in_kind, int64_t n, int64_t pad, float sub, uint16_t mask, float scale,
                    float offset, const CtFilterCoef* coef, int H, int forward_only, float* out, void* workspace,
                    int64_t workspace_bytes, const CtFilterStats* stats, cudaStream_t st);
int ct_filter_forward_seq(const void* in, int in_kind, int64_t n, int64_t pad, float sub, uint16_t mask, float pad_x,
                          const CtFilterCoef* coef, int H, int64_t origin, int part, uint32_t cw_lo, uint32_t cw_step,
                          int64_t cw_begin, int64_t cw_end, uint64_t* counts9, int64_t from_pos, int64_t to_pos,
                          void* workspace, int64_t workspace_bytes, cudaStream_t st);
int ct_filter_backward_seq(int64_t n, int64_t pad, float sub, float scale, float offset, const CtFilterCoef* coef, int H,
                           int64_t origin, float* out, const void* workspace, int64_t workspace_bytes,
                           const CtFilterStats* stats, float* summaries, cudaStream_t st);

extern "C" {

int ct_filter_forward_u16(const uint16_t* raw, int64_t n, int64_t pad, float sub_code, uint16_t mask, float pad_x,
                          const CtFilterCoef* coef, int H, int64_t origin, int part, uint32_t window_lo,
                          uint32_t window_step, int64_t count_begin, int64_t count_end, uint64_t* counts9, int64_t from_pos,
                          int64_t to_pos, void* workspace, int64_t workspace_bytes, void* stream) {
    if (!raw || !coef || n <= 0 || pad < 0 || H < 0 || origin < 0) { ct_set_error("filter_forward: bad argument"); return CT_ERR_ARG; }
    return ct_filter_forward_seq(raw, 0, n, pad, sub_code, mask, pad_x, coef, H, origin, part, window_lo, window_step,
                                 count_begin, count_end, counts9, from_pos, to_pos, workspace, workspace_bytes,
                                 (cudaStream_t)stream);
}

int ct_filter_backward(int64_t n, int64_t pad, float sub_code, float scale, float offset, const CtFilterCoef* coef, int H,
                       int64_t origin, float* out, const void* workspace, int64_t workspace_bytes, const CtFilterStats* stats,
                       float* chunk_minmax, void* stream) {
    if (!out || !coef || n <= 0 || pad < 0 || H < 0 || origin < 0) { ct_set_error("filter_backward: bad argument"); return CT_ERR_ARG; }
    return ct_filter_backward_seq(n, pad, sub_code, scale, offset, coef, H, origin, out, workspace, workspace_bytes, stats,
                                  chunk_minmax, (cudaStream_t)stream);
}

int ct_filter_seq_tile(void) { return 64; }

int ct_filtfilt_u16(const uint16_t* raw, int64_t n, int64_t pad, float median_code, uint16_t mask,
                    float alpha, float pad_value, const CtFilterCoef* coef, int H,
                    int forward_only, float* out, void* workspace, int64_t workspace_bytes, const CtFilterStats* stats,
                    void* stream) {
    if (!raw || !out || !coef || n < 0 || pad < 0 || H < 0) { ct_set_error("filter: bad argument"); return CT_ERR_ARG; }
    if (n == 0) return CT_OK;
    return ct_filtfilt_seq(raw, 0, n, pad, median_code, mask, alpha, pad_value, coef, H, forward_only, out, workspace,
                           workspace_bytes, stats, (cudaStream_t)stream);
}

int ct_filtfilt_f32(const float* x, int64_t n, int64_t pad, float pad_value, const CtFilterCoef* coef,
                    int H, int forward_only, float* out, void* workspace, int64_t workspace_bytes,
                    const CtFilterStats* stats, void* stream) {
    if (!x || !out || !coef || n < 0 || pad < 0 || H < 0) { ct_set_error("filter: bad argument"); return CT_ERR_ARG; }
    if (n == 0) return CT_OK;
    return ct_filtfilt_seq(x, 1, n, pad, pad_value, 0xffff, 1.f, pad_value, coef, H, forward_only, out, workspace,
                           workspace_bytes, stats, (cudaStream_t)stream);
}

int ct_hist_sampled_u16(const uint16_t* raw, int64_t n, int64_t stride, uint16_t mask,
                        uint32_t* hist65536, void* stream) {
    if (!raw || !hist65536 || n < 0 || stride < 1) { ct_set_error("hist: bad argument"); return CT_ERR_ARG; }
    if (n == 0) return CT_OK;
    long long ns = (n + stride - 1) / stride;
    int threads = 256;
    long long blocks = (ns + threads - 1) / threads;
    long long cap = (long long)ct_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    CT_COUNT_LAUNCH();
    ct_hist_sampled_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(raw, n, stride, mask, hist65536);
    return ct_check_launch("ct_hist_sampled_kernel");
}

static int count_window(const uint16_t* raw, int64_t n, uint16_t mask, uint32_t lo, uint32_t step, uint64_t* counts9,
                        bool wide, void* stream) {
    if (!raw || !counts9 || n < 0 || step < 1) { ct_set_error("count_window: bad argument"); return CT_ERR_ARG; }
    if (n == 0) return CT_OK;
    long long blocks = (long long)ct_sm_count() * 8;
    long long want = (n / 8 + 255) / 256 + 1;
    if (blocks > want) blocks = want;
    CT_COUNT_LAUNCH();
    if (wide)
        ct_count_window_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
            raw, n, mask, lo, step, reinterpret_cast<unsigned long long*>(counts9));
    else
        ct_count_window_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
            raw, n, mask, lo, step, reinterpret_cast<unsigned long long*>(counts9));
    return ct_check_launch("ct_count_window_kernel");
}

int ct_count_window_u16(const uint16_t* raw, int64_t n, uint16_t mask, uint32_t lo, uint32_t step,
                        uint64_t* counts9, void* stream) {
    return count_window(raw, n, mask, lo, step, counts9, true, stream);
}

int ct_count_window4_u16(const uint16_t* raw, int64_t n, uint16_t mask, uint32_t lo, uint32_t step,
                         uint64_t* counts9, void* stream) {
    return count_window(raw, n, mask, lo, step, counts9, false, stream);
}

}  // extern "C"
