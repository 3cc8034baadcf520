// Batched Welch periodogram with a hand-written FFT (no cuFFT).
//
// Replaces scipy.signal.welch(x, fs, nperseg=L) as the reference calls it
// (plot-trace.py:442, noise-fit.py:92, legacy/minimal_psd.py:255): periodic Hann window,
// hop L/2, tail dropped, per-segment mean removal, density scaling, one-sided spectrum
// (scipy/signal/_spectral_py.py:515 -> csd -> ShortTimeFFT).
//
// A real segment of L samples (L a power of two) is packed as N = L/2 complex points
// (z[j] = x[2j] + i x[2j+1]), transformed, and split into the real spectrum
// X[k] = (Z[k]+Z*[N-k])/2 - (i/2) e^(-i pi k/N) (Z[k]-Z*[N-k]); |X|^2 is accumulated over
// segments in registers and added once per bin into a float64 accumulator.  Mean removal
// is applied in the spectrum: with the input shifted by a constant c near the mean (for
// float32 headroom), FFT(w (x - mu)) = FFT(w (x - c)) - (mu - c) FFT(w) and FFT(w) of the
// periodic Hann is L/2 at bin 0 and -L/4 at bins +-1.
//
// ONE transform engine (fft_reg): a thread owns 8 points for the whole FFT - radix 8 while
// possible, one radix-4/2 pass last - shared memory carries only the transposes between
// passes (swizzled, double-buffered), no trigonometric table.  Three kernel families use it:
//   * 256 <= L <= 2^14   ct_welch_seg_kernel: a segment is one transform; window, FFT, split
//                        and accumulation without leaving the CTA (4 B/sample of traffic);
//   * 2^15 <= L <= 2^23  four-step N = N1 x N2: ct_welch_cols_fast (window + pack + column
//                        FFTs + twiddle -> Y[k1][n2]) and ct_welch_rows_fast (row FFTs of the
//                        row pair (k1, N1-k1), split, accumulation); the next segment's loads
//                        are in flight while the current one is transformed; batches as large
//                        as the workspace allows (more segments per CTA amortise the set-up);
//   * one segment of ARBITRARY length (nperseg = len(data)): Bluestein chirp-z over generic
//                        complex four-step stages (ct_cfft_cols / ct_cfft_rows).
#include "ct_common.cuh"
#include "cusumtools_b200.h"

namespace {

typedef float2 cpx;
__device__ __forceinline__ cpx cmul(cpx a, cpx b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cpx cadd(cpx a, cpx b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cpx csub(cpx a, cpx b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cpx cconj(cpx a) { return make_float2(a.x, -a.y); }
__device__ __forceinline__ cpx mul_mi(cpx a) { return make_float2(a.y, -a.x); }     // a * (-i)
__device__ __forceinline__ cpx expmi(float t) { float s, c; sincospif(t, &s, &c); return make_float2(c, -s); }  // e^{-i pi t}

template <int R> __device__ __forceinline__ void dft(cpx* v);
template <> __device__ __forceinline__ void dft<4>(cpx* v) {
    cpx a = cadd(v[0], v[2]), b = csub(v[0], v[2]), c = cadd(v[1], v[3]), d = mul_mi(csub(v[1], v[3]));
    v[0] = cadd(a, c); v[1] = cadd(b, d); v[2] = csub(a, c); v[3] = csub(b, d);
}
template <> __device__ __forceinline__ void dft<8>(cpx* v) {
    cpx e[4] = {v[0], v[2], v[4], v[6]}, o[4] = {v[1], v[3], v[5], v[7]};
    dft<4>(e); dft<4>(o);
    const float h = 0.70710678118654752f;
    o[1] = make_float2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));      // * e^{-i pi/4}
    o[2] = mul_mi(o[2]);                                                    // * e^{-i pi/2}
    o[3] = make_float2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));     // * e^{-3i pi/4}
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i] = cadd(e[i], o[i]); v[i + 4] = csub(e[i], o[i]); }
}

struct WelchArgs {
    const float* x; long long n;
    int L, logn1, logn2;       // N = L/2 = 2^logn1 * 2^logn2
    long long seg0; int nseg;  // segments [seg0, seg0+nseg) in this launch
    float c; int use_abs;
    int rsplit;                // the segments of a batch are split over this many CTAs per row pair
    int ssplit;                // ... and over this many CTAs per column group (register-resident kernels)
    cpx* Y;                    // [nseg][N1][N2]
    double* segsum;            // [nseg] sum of (x - c) over the segment
    double* acc;               // [N+1]
    double mu_scale;           // 1/L
};

// =====================================================================================
// The register-resident transform.
//
// The first version of this file kept every Stockham pass in shared memory and took its
// trigonometric factors from an 8 MB table: ncu showed it waiting on L2/HBM latency
// (long_scoreboard 7-11 per issued instruction: one dependent global load per loop iteration,
// scattered table gathers), not on bandwidth or math.  Here a thread owns 8 points of a
// transform for the whole FFT: its 8 global loads are independent and in flight together, the
// first pass reads them straight from registers, the last pass leaves its results in
// registers for the fused epilogue, and shared memory only carries the transposes between
// passes (one write + one read per pass boundary, conflict-free by swizzled rows,
// double-buffered so there is a single barrier per exchange).  No trigonometric table: the
// per-thread twiddle bases are computed once per CTA with sincospif from exactly reduced
// integer arguments, the rest are products with 8th/16th roots of unity and small powers.
// A CTA loops over several segments so the set-up is amortised.
//
// Index algebra (Stockham autosort, radix 8 while possible, one radix-4 or -2 pass last):
// before a pass thread j holds in[j + r N/8], r < 8.  A radix-8 pass with Ns = 8^p finished
// points multiplies input r by e^{-2 pi i (j mod Ns) r / (8 Ns)}, does the 8-point DFT and owns
// out[(j / Ns) 8 Ns + (j mod Ns) + q Ns]; the final radix-R pass (R Ns = N) works on the
// butterflies jj = j + s N/8, s < 8/R, whose inputs are the registers s + (8/R) q and whose
// outputs land in the same registers.  After the last pass thread j holds X[j + r N/8].
// =====================================================================================
constexpr int kFastCols = 8;                 // columns per CTA in the column kernel (64 B of each input row)
constexpr float kH = 0.70710678118654752f;

template <int LOGN> struct FftPlan {
    static constexpr int N = 1 << LOGN;
    static constexpr int T = N / 8;                       // threads per transform
    static constexpr int n8 = LOGN / 3;                   // radix-8 passes
    static constexpr int rem = LOGN % 3;                  // 1: final radix-2 pass, 2: final radix-4 pass
    static constexpr int nbase = n8 - 1 + (rem ? 1 : 0);  // per-thread twiddle bases
    static constexpr int pad = N / 8;                     // swizzle head-room of an exchange buffer (points)
};

// e^{-2 pi i num / den} for integers 0 <= num < den (den a power of two): exact argument reduction
__device__ __forceinline__ cpx root(int num, int den) { return expmi(2.0f * (float)num / (float)den); }

template <int LOGN> __device__ __forceinline__ void fft_bases(cpx* wb, int j) {
    using P = FftPlan<LOGN>;
    int Ns = 8;
#pragma unroll
    for (int p = 1; p < P::n8; ++p) { wb[p - 1] = root(j % Ns, 8 * Ns); Ns *= 8; }
    if (P::rem) wb[P::n8 - 1] = root(j, P::N);
}

// v[r] *= w^r, r = 1..7 (powers by squaring: depth 3)
__device__ __forceinline__ void twiddle8(cpx* v, cpx w) {
    const cpx w2 = cmul(w, w), w4 = cmul(w2, w2), w3 = cmul(w2, w);
    v[1] = cmul(v[1], w); v[2] = cmul(v[2], w2); v[3] = cmul(v[3], w3); v[4] = cmul(v[4], w4);
    v[5] = cmul(v[5], cmul(w4, w)); v[6] = cmul(v[6], cmul(w4, w2)); v[7] = cmul(v[7], cmul(w4, w3));
}

// a * e^{-2 pi i s / 8}
template <int S> __device__ __forceinline__ cpx rot8(cpx a) {
    if (S == 0) return a;
    if (S == 1) return make_float2(kH * (a.x + a.y), kH * (a.y - a.x));
    if (S == 2) return mul_mi(a);
    return make_float2(kH * (a.y - a.x), -kH * (a.x + a.y));           // S == 3
}

// Exchange layouts.  Every exchange is written along out[i0 + q Ns] and read along in[j + r T];
// the swizzles below keep a half-warp's 64-bit accesses on distinct banks AND stay affine along
// both patterns, so each thread needs one base per exchange and the rest are immediates.
//   column kernel (a warp = 4 values of j x 8 columns, rows of 8 points = 16 banks): phi(i) = i + (i >> 3)
//   row kernel (a warp = 32 consecutive j): after pass 0 phi(i) = i + (i >> 4), after pass 1
//   phi(i) = i + 8 (i >> 6) (when N/8 is a multiple of 64), identity afterwards.
template <int LOGN> struct ColsLay {
    using P = FftPlan<LOGN>;
    static constexpr int unit = kFastCols;
    static __device__ __forceinline__ int phi(int, int i) { return i + (i >> 3); }
    static __host__ __device__ constexpr int wstride(int p, int Ns) { return p == 0 ? 1 : Ns + Ns / 8; }
    static __host__ __device__ constexpr int rstride(int) { return P::T + P::T / 8; }
};
template <int LOGN> struct RowsLay {
    using P = FftPlan<LOGN>;
    static constexpr int unit = 1;
    static constexpr bool sw1 = (P::T % 64) == 0;
    static __device__ __forceinline__ int phi(int p, int i) {
        return p == 0 ? i + (i >> 4) : ((p == 1 && sw1) ? i + ((i >> 6) << 3) : i);
    }
    static __host__ __device__ constexpr int wstride(int p, int Ns) { return p == 0 ? 1 : Ns; }
    static __host__ __device__ constexpr int rstride(int p) { return p == 0 ? P::T + P::T / 16 : ((p == 1 && sw1) ? P::T + P::T / 8 : P::T); }
};
template <int LOGN> struct FftX {                       // exchanges inside one transform
    static constexpr int n = FftPlan<LOGN>::n8 - 1 + (FftPlan<LOGN>::rem ? 1 : 0);
};

// per-thread exchange bases (in points, before the `unit` scaling and the slot/column offset)
template <int LOGN, class Lay> __device__ __forceinline__ void fft_exchange_bases(int* wbs, int* rbs, int j, int off) {
    int Ns = 1;
#pragma unroll
    for (int p = 0; p < FftX<LOGN>::n; ++p) {
        const int i0 = (j / Ns) * Ns * 8 + (j % Ns);
        wbs[p] = Lay::phi(p, i0) * Lay::unit + off;
        rbs[p] = Lay::phi(p, j) * Lay::unit + off;
        Ns *= 8;
    }
}

// Full transform of the 8 points in v (see the index algebra above).  Exchange e uses buffer
// e & 1 (callers keep the number of exchanges per loop iteration even, or add a barrier).
template <int LOGN, class Lay>
__device__ __forceinline__ void fft_reg(cpx (&v)[8], const cpx* wb, cpx* buf0, cpx* buf1, const int* wbs, const int* rbs) {
    using P = FftPlan<LOGN>;
    int Ns = 1;
#pragma unroll
    for (int p = 0; p < P::n8; ++p) {
        if (p > 0) twiddle8(v, wb[p - 1]);
        dft<8>(v);
        if (p + 1 < P::n8 || P::rem) {
            cpx* b = (p & 1) ? buf1 : buf0;
            cpx* w = b + wbs[p];
#pragma unroll
            for (int q = 0; q < 8; ++q) w[q * Lay::wstride(p, Ns) * Lay::unit] = v[q];
            __syncthreads();
            const cpx* rd = b + rbs[p];
#pragma unroll
            for (int r = 0; r < 8; ++r) v[r] = rd[r * Lay::rstride(p) * Lay::unit];
        }
        Ns *= 8;
    }
    if (P::rem == 2) {
        const cpx w0 = wb[P::n8 - 1], w1 = rot8<1>(w0);
        {   cpx t[4] = {v[0], v[2], v[4], v[6]};
            const cpx w2 = cmul(w0, w0);
            t[1] = cmul(t[1], w0); t[2] = cmul(t[2], w2); t[3] = cmul(t[3], cmul(w2, w0));
            dft<4>(t); v[0] = t[0]; v[2] = t[1]; v[4] = t[2]; v[6] = t[3]; }
        {   cpx t[4] = {v[1], v[3], v[5], v[7]};
            const cpx w2 = cmul(w1, w1);
            t[1] = cmul(t[1], w1); t[2] = cmul(t[2], w2); t[3] = cmul(t[3], cmul(w2, w1));
            dft<4>(t); v[1] = t[0]; v[3] = t[1]; v[5] = t[2]; v[7] = t[3]; }
    } else if (P::rem == 1) {
        const cpx w0 = wb[P::n8 - 1];
        const cpx ws[4] = {w0, rot8<1>(w0), rot8<2>(w0), rot8<3>(w0)};
#pragma unroll
        for (int s2 = 0; s2 < 4; ++s2) {
            const cpx a = v[s2], b = cmul(v[s2 + 4], ws[s2]);
            v[s2] = cadd(a, b); v[s2 + 4] = csub(a, b);
        }
    }
}

// Column kernel: window + pack + length-N1 FFT down kFastCols adjacent columns + four-step twiddle.
// Thread (j, c): column n2 = 8 g + c, points n1 = j + r N1/8.  The next segment's 8 loads are issued
// before the current segment is transformed.
template <int LOG1>
__global__ void __launch_bounds__((1 << LOG1) / 8 * kFastCols) ct_welch_cols_fast(WelchArgs a) {
    using P = FftPlan<LOG1>;
    using Lay = ColsLay<LOG1>;
    constexpr int N1 = P::N, T1 = P::T, NX = FftX<LOG1>::n;
    extern __shared__ __align__(16) unsigned char smraw[];
    cpx* buf0 = reinterpret_cast<cpx*>(smraw);
    cpx* buf1 = buf0 + (size_t)(N1 + P::pad) * kFastCols;
    __shared__ double wsum[2][32];
    const int N2 = 1 << a.logn2;
    const long long N = (long long)N1 * N2;
    const int tid = threadIdx.x, c = tid % kFastCols, j = tid / kFastCols;
    const int groups = N2 / kFastCols;
    const int g = blockIdx.x % groups, sp = blockIdx.x / groups;
    const int n2 = g * kFastCols + c;
    // ---- per-thread constants
    cpx wb[P::nbase > 0 ? P::nbase : 1];
    fft_bases<LOG1>(wb, j);
    int wbs[NX > 0 ? NX : 1], rbs[NX > 0 ? NX : 1];
    fft_exchange_bases<LOG1, Lay>(wbs, rbs, j, c);
    // periodic Hann at the samples 2J, 2J+1 of point J = N2 (j + r T1) + n2: the angle advances by pi/4 per r
    const long long J0 = (long long)N2 * j + n2;
    float ce, se, co, so;
    sincospif(2.0f * (float)(2 * J0) / (float)a.L, &se, &ce);
    sincospif(2.0f * (float)(2 * J0 + 1) / (float)a.L, &so, &co);
    // four-step twiddle W_N^(n2 k1), k1 = j + r T1
    cpx ot[8];
    {
        const cpx base = root((int)(((long long)n2 * j) % N), (int)N), step = root((int)(((long long)n2 * T1) % N), (int)N);
        const cpx s2 = cmul(step, step), s4 = cmul(s2, s2);
        ot[0] = base; ot[1] = cmul(base, step); ot[2] = cmul(base, s2); ot[3] = cmul(ot[1], s2);
        ot[4] = cmul(base, s4); ot[5] = cmul(ot[1], s4); ot[6] = cmul(ot[2], s4); ot[7] = cmul(ot[3], s4);
    }
    const long long rstep = 2 * (long long)N2 * T1;        // floats between a thread's consecutive points
    const float* x0p = a.x + a.seg0 * (long long)(a.L / 2) + 2 * J0;
    float2 in[8], nx[8];
    if (sp < a.nseg) {
        const float* xs = x0p + sp * (long long)(a.L / 2);
#pragma unroll
        for (int r = 0; r < 8; ++r) in[r] = __ldg(reinterpret_cast<const float2*>(xs + rstep * r));
    }
    for (int s = sp, it = 0; s < a.nseg; s += a.ssplit, ++it) {
        if (s + a.ssplit < a.nseg) {
            const float* xs = x0p + (s + a.ssplit) * (long long)(a.L / 2);
#pragma unroll
            for (int r = 0; r < 8; ++r) nx[r] = __ldg(reinterpret_cast<const float2*>(xs + rstep * r));
        }
        cpx v[8];
        float part = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            float x0 = in[r].x, x1 = in[r].y;
            if (a.use_abs) { x0 = fabsf(x0); x1 = fabsf(x1); }
            x0 -= a.c; x1 -= a.c;
            part += x0 + x1;
            // cos(theta0 + r pi/4) for the even and the odd sample
            float cer, cor;
            switch (r) {
                case 0: cer = ce; cor = co; break;
                case 1: cer = kH * (ce - se); cor = kH * (co - so); break;
                case 2: cer = -se; cor = -so; break;
                case 3: cer = -kH * (ce + se); cor = -kH * (co + so); break;
                case 4: cer = -ce; cor = -co; break;
                case 5: cer = -kH * (ce - se); cor = -kH * (co - so); break;
                case 6: cer = se; cor = so; break;
                default: cer = kH * (ce + se); cor = kH * (co + so); break;
            }
            v[r] = make_float2(x0 * (0.5f - 0.5f * cer), x1 * (0.5f - 0.5f * cor));
        }
        // segment sum for the mean: warp, CTA, one atomic per CTA and segment.  The partial sums ride
        // on the FFT's barriers (written before its first exchange, read after its last); two copies
        // because a fast warp may already be one segment ahead.
        double dp = (double)part;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dp += __shfl_xor_sync(CT_FULL, dp, o);
        double* wsm = wsum[it & 1];
        if ((tid & 31) == 0) wsm[tid >> 5] = dp;
        fft_reg<LOG1, Lay>(v, wb, buf0, buf1, wbs, rbs);
        if (NX & 1) __syncthreads();                       // keep the buffer parity of the next iteration safe
        if (tid == 0) {
            double t = 0;
            for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += wsm[i];
            atomicAdd(a.segsum + s, t);
        }
        cpx* Y = a.Y + (size_t)s * N + n2;
#pragma unroll
        for (int r = 0; r < 8; ++r) Y[(size_t)(j + r * T1) * N2] = cmul(v[r], ot[r]);
#pragma unroll
        for (int r = 0; r < 8; ++r) in[r] = nx[r];
    }
}

// Row kernel: length-N2 FFT along the row pair (k1, N1 - k1), real-FFT split, |X|^2 accumulated over the
// CTA's segments in registers.  Thread (p, j): row p of the pair, bins k2 = j + r N2/8.
template <int LOG2>
__global__ void __launch_bounds__(2 * (1 << LOG2) / 8) ct_welch_rows_fast(WelchArgs a) {
    using P = FftPlan<LOG2>;
    using Lay = RowsLay<LOG2>;
    constexpr int N2 = P::N, T2 = P::T, kSlot = N2 + P::pad, NX = FftX<LOG2>::n;
    extern __shared__ __align__(16) unsigned char smraw[];
    cpx* buf0 = reinterpret_cast<cpx*>(smraw);
    cpx* buf1 = buf0 + 2 * kSlot;
    const int N1 = 1 << a.logn1;
    const long long N = (long long)N1 * N2;
    const int tid = threadIdx.x, p = tid / T2, j = tid % T2;
    const int r0 = blockIdx.x / a.rsplit, part = blockIdx.x % a.rsplit;
    const int r1 = (r0 == 0) ? 0 : N1 - r0;        // partner row (== r0 for rows 0 and N1/2)
    const bool self = (r1 == r0);
    const int row = p ? r1 : r0;
    cpx wb[P::nbase > 0 ? P::nbase : 1];
    fft_bases<LOG2>(wb, j);
    int wbs[NX > 0 ? NX : 1], rbs[NX > 0 ? NX : 1];
    fft_exchange_bases<LOG2, Lay>(wbs, rbs, j, p * kSlot);
    // split twiddle e^{-2 pi i k / L} at k = row + N1 (j + r T2): the factor per r is a 16th root of unity
    const cpx tb = root(row + N1 * j, a.L);
    const int mine = p * kSlot + j;
    // partner bin of k2 = j + r T2: N2 - 1 - k2 in the other row; (N2 - k2) mod N2 in the same row for row 0
    const int other = (1 - p) * kSlot + N2 - (r0 == 0 ? 0 : 1) - j;
    const bool owner = !(self && p == 1);          // the second slot of a self-paired row only mirrors the first
    const bool nyq = self && p == 1 && r0 == 0 && j == 0;
    float acc[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = 0.f;
    const float hl = 0.5f * (float)a.L, ql = 0.25f * (float)a.L;
    const cpx* Y0 = a.Y + (size_t)row * N2 + j;
    cpx in[8], nx[8];
    if (part < a.nseg) {
#pragma unroll
        for (int r = 0; r < 8; ++r) in[r] = Y0[(size_t)part * N + r * T2];
    }
    for (int s = part; s < a.nseg; s += a.rsplit) {
        if (s + a.rsplit < a.nseg) {
#pragma unroll
            for (int r = 0; r < 8; ++r) nx[r] = Y0[(size_t)(s + a.rsplit) * N + r * T2];
        }
        cpx v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) v[r] = in[r];
        const float dmu = (float)(a.segsum[s] * a.mu_scale);     // mu - c for this segment
        fft_reg<LOG2, Lay>(v, wb, buf0, buf1, wbs, rbs);
        cpx* b = (NX & 1) ? buf1 : buf0;
#pragma unroll
        for (int r = 0; r < 8; ++r) b[mine + r * T2] = v[r];
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            int mi = other - r * T2;
            if (r == 0 && r0 == 0 && j == 0) mi -= N2;            // (N2 - 0) mod N2
            const cpx Zk = v[r], Zm = b[mi];
            const cpx E = cadd(Zk, cconj(Zm)), O = csub(Zk, cconj(Zm));
            // tw = tb * e^{-2 pi i r / 16}
            cpx tw;
            {
                const float c16 = 0.92387953251128674f, s16 = 0.38268343236508977f;
                const cpx h = (r & 1) ? make_float2(tb.x * c16 + tb.y * s16, tb.y * c16 - tb.x * s16) : tb;   // * e^{-i pi/8}
                switch (r >> 1) { case 0: tw = h; break; case 1: tw = rot8<1>(h); break; case 2: tw = rot8<2>(h); break; default: tw = rot8<3>(h); break; }
            }
            cpx X = cadd(make_float2(0.5f * E.x, 0.5f * E.y), cmul(make_float2(0.5f * O.y, -0.5f * O.x), tw));
            if (r == 0) {
                const long long k = row + (long long)N1 * j;
                if (k == 0) X.x -= dmu * hl;
                if (k == 1) X.x += dmu * ql;
                if (nyq) X = make_float2(Zk.x - Zk.y, 0.f);      // bin N lives in Z[0]
            }
            acc[r] += X.x * X.x + X.y * X.y;
        }
        if (!(NX & 1)) __syncthreads();                           // odd number of exchanges per segment
#pragma unroll
        for (int r = 0; r < 8; ++r) in[r] = nx[r];
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const long long k = row + (long long)N1 * (j + r * T2);
        if (owner) atomicAdd(a.acc + k, (double)acc[r]);
        else if (nyq && r == 0) atomicAdd(a.acc + N, (double)acc[r]);
    }
}

// Whole-segment kernel (256 <= L <= 2^14): the N = L/2 packed points of a segment fit one transform, so a
// segment is windowed, transformed, split and accumulated without leaving the CTA.  A CTA holds S slots
// (S T >= 128 threads for the short lengths) and walks the segments S at a time; slot p, thread j owns the
// points / bins j + r N/8.
template <int LOGN>
__global__ void __launch_bounds__((1 << LOGN) / 8 >= 128 ? (1 << LOGN) / 8 : 128) ct_welch_seg_kernel(WelchArgs a) {
    using P = FftPlan<LOGN>;
    using Lay = RowsLay<LOGN>;
    constexpr int N = P::N, T = P::T, kSlot = N + P::pad, NX = FftX<LOGN>::n;
    constexpr int S = T >= 128 ? 1 : 128 / T;                 // slots per CTA
    constexpr int WPS = T >= 32 ? T / 32 : 1;                 // warps per slot
    extern __shared__ __align__(16) unsigned char smraw[];
    cpx* buf0 = reinterpret_cast<cpx*>(smraw);
    cpx* buf1 = buf0 + S * kSlot;
    __shared__ float wsum[2][S * WPS];
    const int tid = threadIdx.x, p = tid / T, j = tid % T;
    const int lane = ct_lane();
    cpx wb[P::nbase > 0 ? P::nbase : 1];
    fft_bases<LOGN>(wb, j);
    int wbs[NX > 0 ? NX : 1], rbs[NX > 0 ? NX : 1];
    fft_exchange_bases<LOGN, Lay>(wbs, rbs, j, p * kSlot);
    // periodic Hann at samples 2J, 2J+1, J = j + r T: the angle advances by pi/4 per r
    float ce, se, co, so;
    sincospif(2.0f * (float)(2 * j) / (float)a.L, &se, &ce);
    sincospif(2.0f * (float)(2 * j + 1) / (float)a.L, &so, &co);
    const cpx tb = root(j, a.L);                              // split twiddle e^{-2 pi i k / L} at k = j (+ r T: 16th roots)
    const int mine = p * kSlot + j;
    const int other = p * kSlot + N - j;                      // partner bin (N - k) mod N of k = j + r T
    float acc[8], accn = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = 0.f;
    const float hl = 0.5f * (float)a.L, ql = 0.25f * (float)a.L;
    const int stride = gridDim.x * S;
    const int iters = (a.nseg + stride - 1) / stride;         // every slot runs the same number of barriers
    for (int it = 0; it < iters; ++it) {
        const int sg = blockIdx.x * S + p + it * stride;
        const bool live = sg < a.nseg;
        cpx v[8];
        float part = 0.f;
        const float* xs = a.x + (a.seg0 + (live ? sg : 0)) * (long long)(a.L / 2) + 2 * j;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            float2 in = make_float2(a.c, a.c);
            if (live) in = __ldg(reinterpret_cast<const float2*>(xs + 2 * T * r));
            float x0 = in.x, x1 = in.y;
            if (a.use_abs) { x0 = fabsf(x0); x1 = fabsf(x1); }
            x0 -= a.c; x1 -= a.c;
            part += x0 + x1;
            float cer, cor;
            switch (r) {
                case 0: cer = ce; cor = co; break;
                case 1: cer = kH * (ce - se); cor = kH * (co - so); break;
                case 2: cer = -se; cor = -so; break;
                case 3: cer = -kH * (ce + se); cor = -kH * (co + so); break;
                case 4: cer = -ce; cor = -co; break;
                case 5: cer = -kH * (ce - se); cor = -kH * (co - so); break;
                case 6: cer = se; cor = so; break;
                default: cer = kH * (ce + se); cor = kH * (co + so); break;
            }
            v[r] = make_float2(x0 * (0.5f - 0.5f * cer), x1 * (0.5f - 0.5f * cor));
        }
        // segment sum: shuffles inside the slot's lanes, then across its warps through shared memory
        // (written before the transform's first barrier, read after its last; two copies per parity)
#pragma unroll
        for (int o = (T < 32 ? T : 32) / 2; o > 0; o >>= 1) part += __shfl_xor_sync(CT_FULL, part, o);
        float* wsm = wsum[it & 1];
        if (T >= 32) { if (lane == 0) wsm[p * WPS + (j >> 5)] = part; }
        fft_reg<LOGN, Lay>(v, wb, buf0, buf1, wbs, rbs);
        cpx* b = (NX & 1) ? buf1 : buf0;
#pragma unroll
        for (int r = 0; r < 8; ++r) b[mine + r * T] = v[r];
        __syncthreads();
        float tot = part;
        if (T >= 32) {
            tot = 0.f;
#pragma unroll
            for (int w = 0; w < WPS; ++w) tot += wsm[p * WPS + w];
        }
        const float dmu = tot / (float)a.L;                   // mu - c of this segment
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            int mi = other - r * T;
            if (r == 0 && j == 0) mi -= N;
            const cpx Zk = v[r], Zm = b[mi];
            const cpx E = cadd(Zk, cconj(Zm)), O = csub(Zk, cconj(Zm));
            cpx tw;
            {
                const float c16 = 0.92387953251128674f, s16 = 0.38268343236508977f;
                const cpx h = (r & 1) ? make_float2(tb.x * c16 + tb.y * s16, tb.y * c16 - tb.x * s16) : tb;
                switch (r >> 1) { case 0: tw = h; break; case 1: tw = rot8<1>(h); break; case 2: tw = rot8<2>(h); break; default: tw = rot8<3>(h); break; }
            }
            cpx X = cadd(make_float2(0.5f * E.x, 0.5f * E.y), cmul(make_float2(0.5f * O.y, -0.5f * O.x), tw));
            if (r == 0) {
                if (j == 0) {
                    X.x -= dmu * hl;                                        // bin 0
                    if (live) accn += (Zk.x - Zk.y) * (Zk.x - Zk.y);       // bin N (Nyquist) lives in Z[0]
                }
                if (j == 1) X.x += dmu * ql;                               // bin 1
            }
            if (live) acc[r] += X.x * X.x + X.y * X.y;
        }
        if (!(NX & 1)) __syncthreads();                       // odd number of exchanges per iteration
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) atomicAdd(a.acc + j + r * T, (double)acc[r]);
    if (j == 0) atomicAdd(a.acc + N, (double)accn);
}

template <int LOGN> static int launch_seg(const WelchArgs& a, cudaStream_t st) {
    using P = FftPlan<LOGN>;
    constexpr int S = P::T >= 128 ? 1 : 128 / P::T;
    const size_t sm = (size_t)2 * S * (P::N + P::pad) * sizeof(cpx);
    cudaFuncSetAttribute(ct_welch_seg_kernel<LOGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ct_welch_seg_kernel<LOGN>, S * P::T, sm);
    if (occ < 1) occ = 1;
    long long grid = (long long)ct_sm_count() * occ;
    const long long want = (a.nseg + S - 1) / S;
    if (grid > want) grid = want;
    CT_COUNT_LAUNCH();
    ct_welch_seg_kernel<LOGN><<<(unsigned)grid, S * P::T, sm, st>>>(a);
    return ct_check_launch("ct_welch_seg_kernel");
}
static int seg_path(const WelchArgs& a, int logn, cudaStream_t st) {
    switch (logn) {
        case 7: return launch_seg<7>(a, st);
        case 8: return launch_seg<8>(a, st);
        case 9: return launch_seg<9>(a, st);
        case 10: return launch_seg<10>(a, st);
        case 11: return launch_seg<11>(a, st);
        case 12: return launch_seg<12>(a, st);
        default: return launch_seg<13>(a, st);
    }
}

template <int LOG1> static int launch_cols_fast(const WelchArgs& a, int N2, cudaStream_t st) {
    using P = FftPlan<LOG1>;
    const size_t sm = (size_t)2 * (P::N + P::pad) * kFastCols * sizeof(cpx);
    cudaFuncSetAttribute(ct_welch_cols_fast<LOG1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    CT_COUNT_LAUNCH();
    ct_welch_cols_fast<LOG1><<<(unsigned)((N2 / kFastCols) * a.ssplit), P::T * kFastCols, sm, st>>>(a);
    return ct_check_launch("ct_welch_cols_fast");
}
template <int LOG2> static int launch_rows_fast(const WelchArgs& a, int N1, cudaStream_t st) {
    using P = FftPlan<LOG2>;
    const size_t sm = (size_t)2 * 2 * (P::N + P::pad) * sizeof(cpx);
    cudaFuncSetAttribute(ct_welch_rows_fast<LOG2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    CT_COUNT_LAUNCH();
    ct_welch_rows_fast<LOG2><<<(unsigned)((N1 / 2 + 1) * a.rsplit), 2 * P::T, sm, st>>>(a);
    return ct_check_launch("ct_welch_rows_fast");
}
static int cols_fast(const WelchArgs& a, cudaStream_t st) {
    const int N2 = 1 << a.logn2;
    switch (a.logn1) {
        case 7: return launch_cols_fast<7>(a, N2, st);
        case 8: return launch_cols_fast<8>(a, N2, st);
        case 9: return launch_cols_fast<9>(a, N2, st);
        default: return launch_cols_fast<10>(a, N2, st);
    }
}
static int rows_fast(const WelchArgs& a, cudaStream_t st) {
    const int N1 = 1 << a.logn1;
    switch (a.logn2) {
        case 7: return launch_rows_fast<7>(a, N1, st);
        case 8: return launch_rows_fast<8>(a, N1, st);
        case 9: return launch_rows_fast<9>(a, N1, st);
        case 10: return launch_rows_fast<10>(a, N1, st);
        case 11: return launch_rows_fast<11>(a, N1, st);
        default: return launch_rows_fast<12>(a, N1, st);
    }
}


// =====================================================================================
// One segment of ARBITRARY length (plot-trace.py:437: nperseg = min(2^20, len) is the window
// length itself whenever the displayed window is shorter than 2^20 samples, and the psd-length
// branch caps at len, :433-435): the length-n DFT by Bluestein's chirp-z identity
//     X[k] = c_k * sum_j (x_j c_j) conj(c)_(k-j),   c_m = e^{-i pi m^2 / n},
// i.e. one circular convolution of length M = 2^m >= 2n - 1 = three power-of-two complex FFTs
// built from the same register-resident transform (four-step N1 x N2, generic load / store
// stages).  |c_k| = 1, so the power spectrum needs no final chirp.  Not a hot path: one
// segment, n < 2^21, once per call.
// =====================================================================================
struct CfftArgs {
    const cpx* in; const cpx* mul; long long n_valid; int conj_in;   // input (zero beyond n_valid), optional pointwise factor
    cpx* Y;                                                           // four-step intermediate [N1][N2]
    cpx* out; int conj_out; float scale;                              // natural-order result
    int logn1, logn2;
};

template <int LOG1>
__global__ void __launch_bounds__((1 << LOG1) / 8 * kFastCols) ct_cfft_cols(CfftArgs a) {
    using P = FftPlan<LOG1>;
    using Lay = ColsLay<LOG1>;
    constexpr int N1 = P::N, T1 = P::T, NX = FftX<LOG1>::n;
    extern __shared__ __align__(16) unsigned char smraw[];
    cpx* buf0 = reinterpret_cast<cpx*>(smraw);
    cpx* buf1 = buf0 + (size_t)(N1 + P::pad) * kFastCols;
    const int N2 = 1 << a.logn2;
    const long long N = (long long)N1 * N2;
    const int tid = threadIdx.x, c = tid % kFastCols, j = tid / kFastCols;
    const int n2 = blockIdx.x * kFastCols + c;
    cpx wb[P::nbase > 0 ? P::nbase : 1];
    fft_bases<LOG1>(wb, j);
    int wbs[NX > 0 ? NX : 1], rbs[NX > 0 ? NX : 1];
    fft_exchange_bases<LOG1, Lay>(wbs, rbs, j, c);
    cpx v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const long long J = (long long)N2 * (j + r * T1) + n2;
        cpx x = make_float2(0.f, 0.f);
        if (J < a.n_valid) {
            x = a.in[J];
            if (a.mul) x = cmul(x, a.mul[J]);
            if (a.conj_in) x = cconj(x);
        }
        v[r] = x;
    }
    fft_reg<LOG1, Lay>(v, wb, buf0, buf1, wbs, rbs);
    const cpx base = root((int)(((long long)n2 * j) % N), (int)N), step = root((int)(((long long)n2 * T1) % N), (int)N);
    cpx tw = base;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        a.Y[(size_t)(j + r * T1) * N2 + n2] = cmul(v[r], tw);
        tw = cmul(tw, step);
    }
}

template <int LOG2>
__global__ void __launch_bounds__(2 * (1 << LOG2) / 8) ct_cfft_rows(CfftArgs a) {
    using P = FftPlan<LOG2>;
    using Lay = RowsLay<LOG2>;
    constexpr int N2 = P::N, T2 = P::T, kSlot = N2 + P::pad, NX = FftX<LOG2>::n;
    extern __shared__ __align__(16) unsigned char smraw[];
    cpx* buf0 = reinterpret_cast<cpx*>(smraw);
    cpx* buf1 = buf0 + 2 * kSlot;
    const int N1 = 1 << a.logn1;
    const int tid = threadIdx.x, p = tid / T2, j = tid % T2;
    const int row = 2 * blockIdx.x + p;                      // N1 is even: two rows per CTA
    cpx wb[P::nbase > 0 ? P::nbase : 1];
    fft_bases<LOG2>(wb, j);
    int wbs[NX > 0 ? NX : 1], rbs[NX > 0 ? NX : 1];
    fft_exchange_bases<LOG2, Lay>(wbs, rbs, j, p * kSlot);
    cpx v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = a.Y[(size_t)row * N2 + j + r * T2];
    fft_reg<LOG2, Lay>(v, wb, buf0, buf1, wbs, rbs);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        cpx z = make_float2(v[r].x * a.scale, v[r].y * a.scale);
        if (a.conj_out) z = cconj(z);
        a.out[row + (size_t)N1 * (j + r * T2)] = z;
    }
}

template <int LOG1> static int launch_cfft_cols(const CfftArgs& a, cudaStream_t st) {
    using P = FftPlan<LOG1>;
    const size_t sm = (size_t)2 * (P::N + P::pad) * kFastCols * sizeof(cpx);
    cudaFuncSetAttribute(ct_cfft_cols<LOG1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    CT_COUNT_LAUNCH();
    ct_cfft_cols<LOG1><<<(unsigned)((1 << a.logn2) / kFastCols), P::T * kFastCols, sm, st>>>(a);
    return ct_check_launch("ct_cfft_cols");
}
template <int LOG2> static int launch_cfft_rows(const CfftArgs& a, cudaStream_t st) {
    using P = FftPlan<LOG2>;
    const size_t sm = (size_t)2 * 2 * (P::N + P::pad) * sizeof(cpx);
    cudaFuncSetAttribute(ct_cfft_rows<LOG2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    CT_COUNT_LAUNCH();
    ct_cfft_rows<LOG2><<<(unsigned)((1 << a.logn1) / 2), 2 * P::T, sm, st>>>(a);
    return ct_check_launch("ct_cfft_rows");
}
// complex FFT of M = 2^(logn1+logn2) points, 6 <= logn1 <= 10, 7 <= logn2 <= 12
static int cfft(const CfftArgs& a, cudaStream_t st) {
    int rc;
    switch (a.logn1) {
        case 6: rc = launch_cfft_cols<6>(a, st); break;
        case 7: rc = launch_cfft_cols<7>(a, st); break;
        case 8: rc = launch_cfft_cols<8>(a, st); break;
        case 9: rc = launch_cfft_cols<9>(a, st); break;
        default: rc = launch_cfft_cols<10>(a, st); break;
    }
    if (rc) return rc;
    switch (a.logn2) {
        case 7: return launch_cfft_rows<7>(a, st);
        case 8: return launch_cfft_rows<8>(a, st);
        case 9: return launch_cfft_rows<9>(a, st);
        case 10: return launch_cfft_rows<10>(a, st);
        case 11: return launch_cfft_rows<11>(a, st);
        default: return launch_cfft_rows<12>(a, st);
    }
}

// c[j] = e^{-i pi j^2 / n}: j^2 reduced mod 2n in integers, the angle in float64
__global__ void ct_bluestein_chirp(cpx* __restrict__ c, long long n) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const unsigned long long r = ((unsigned long long)j * (unsigned long long)j) % (unsigned long long)(2 * n);
    double sn, cs;
    sincospi((double)r / (double)n, &sn, &cs);
    c[j] = make_float2((float)cs, (float)(-sn));
}
// the convolution kernel conj(c)_(m) wrapped to length M: b[m] = conj(c[|m|]) for |m| < n (indices mod M), else 0
__global__ void ct_bluestein_kernel(cpx* __restrict__ b, const cpx* __restrict__ c, long long n, long long M) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    cpx v = make_float2(0.f, 0.f);
    if (m < n) v = cconj(c[m]);
    else if (M - m < n) v = cconj(c[M - m]);
    b[m] = v;
}
// a[j] = hann_n(j) (x_j - mean) c[j]   (periodic Hann; |x| first if use_abs)
__global__ void ct_bluestein_input(cpx* __restrict__ a, const float* __restrict__ x, const cpx* __restrict__ c, long long n,
                                   double mean, int use_abs) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double v = (double)x[j];
    if (use_abs) v = fabs(v);
    const double w = 0.5 - 0.5 * cospi(2.0 * (double)j / (double)n);
    const float s = (float)(w * (v - mean));
    a[j] = make_float2(s * c[j].x, s * c[j].y);
}
__global__ void ct_bluestein_power(double* __restrict__ acc, const cpx* __restrict__ z, long long nout) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nout) return;
    const double re = (double)z[k].x, im = (double)z[k].y;
    acc[k] = re * re + im * im;
}
static int bluestein_logm(long long n) {
    int m = 13;
    while ((1LL << m) < 2 * n - 1) ++m;
    return m;
}

}  // namespace

extern "C" {

// CTAs of `threads` threads resident on the GPU at once (the register-resident kernels use up to 128 registers)
static int resident_ctas(int threads) {
    int per_sm = 65536 / (128 * threads);
    if (per_sm < 1) per_sm = 1;
    return ct_sm_count() * per_sm;
}
// Number of CTAs the segments of a batch are split over per unit (row pair / column group).
static int pick_split(int units, int nseg, int resident) {
    int best = 1;
    double best_cost = 1e300;
    for (int s = 1; s <= nseg && s <= 64; ++s) {
        const long long grid = (long long)units * s;
        const double waves = (double)((grid + resident - 1) / resident);
        const double cost = waves * ((double)((nseg + s - 1) / s) + 0.5);
        if (cost < best_cost * (1.0 - 1e-9)) { best_cost = cost; best = s; }
    }
    return best;
}

int64_t ct_welch_workspace_bytes(int32_t nperseg, int32_t batch) {
    if (nperseg < 256 || (nperseg & (nperseg - 1))) return -1;
    if (nperseg <= (1 << 14)) return 256;                      // whole-segment kernel: no intermediate
    return (int64_t)batch * (nperseg / 2) * 8 + (int64_t)batch * 8 + 256;
}

int ct_welch_f32(const float* x, int64_t n, int32_t nperseg, float shift, int32_t use_abs, int32_t batch,
                 void* workspace, int64_t workspace_bytes, double* acc, int64_t* nseg_out, void* stream) {
    if (!x || !workspace || !acc || !nseg_out) { ct_set_error("welch: null pointer"); return CT_ERR_ARG; }
    if (nperseg < 256 || (nperseg & (nperseg - 1)) || nperseg > (1 << 23)) {
        ct_set_error("welch: nperseg must be a power of two in [256, 2^23] (got %d)", nperseg); return CT_ERR_UNSUPPORTED;
    }
    if ((reinterpret_cast<uintptr_t>(x) & 7) != 0) { ct_set_error("welch: input must be 8-byte aligned"); return CT_ERR_ARG; }
    const int L = nperseg, hop = L / 2;
    const long long nseg = n >= L ? (n - (L - hop)) / hop : 0;
    *nseg_out = nseg;
    cudaStream_t st = (cudaStream_t)stream;
    const long long N = L / 2;
    cudaMemsetAsync(acc, 0, (N + 1) * sizeof(double), st);
    if (nseg == 0) return CT_OK;
    if (batch < 1) batch = 1;
    if (workspace_bytes < ct_welch_workspace_bytes(nperseg, batch)) { ct_set_error("welch: workspace too small"); return CT_ERR_ARG; }
    int logn = 0; while ((1LL << logn) < N) ++logn;
    WelchArgs a;
    a.x = x; a.n = n; a.L = L; a.c = shift; a.use_abs = use_abs; a.acc = acc; a.mu_scale = 1.0 / (double)L;
    a.rsplit = a.ssplit = 1; a.logn1 = a.logn2 = 0; a.Y = nullptr; a.segsum = nullptr;
    if (logn <= 13) {                                          // L <= 2^14: a segment is one transform, no intermediate
        if (nseg > 0x7fffffffLL) { ct_set_error("welch: too many segments"); return CT_ERR_UNSUPPORTED; }
        a.seg0 = 0; a.nseg = (int)nseg;
        return seg_path(a, logn, st);
    }
    // four-step N1 x N2: N2 >= N1, the column kernel holds 8 columns of N1 <= 1024 points
    int logn2 = (logn + 1) / 2, logn1 = logn - logn2;
    if (logn1 > 10) { logn1 = 10; logn2 = logn - 10; }
    a.logn1 = logn1; a.logn2 = logn2;
    a.Y = (cpx*)workspace;
    a.segsum = (double*)((char*)workspace + (size_t)batch * N * 8);
    const int N1 = 1 << logn1, N2 = 1 << logn2;
    for (long long s0 = 0; s0 < nseg; s0 += batch) {
        a.seg0 = s0; a.nseg = (int)((nseg - s0 < batch) ? nseg - s0 : batch);
        cudaMemsetAsync(a.segsum, 0, (size_t)a.nseg * 8, st);
        // enough CTAs to fill the GPU: the segments of the batch are split over rsplit CTAs per row pair
        // (and over ssplit CTAs per column group); more segments per CTA amortise its set-up
        // (chosen against the wave quantisation: the CTAs of a launch run in ceil(grid / resident) waves, and a CTA
        // costs its segments plus about half a segment of set-up)
        a.rsplit = pick_split(N1 / 2 + 1, a.nseg, resident_ctas(2 * (N2 / 8)));
        a.ssplit = pick_split(N2 / kFastCols, a.nseg, resident_ctas((N1 / 8) * kFastCols));
        int rc = cols_fast(a, st); if (rc) return rc;
        rc = rows_fast(a, st); if (rc) return rc;
    }
    return CT_OK;
}

int64_t ct_welch_single_workspace_bytes(int64_t n) {
    if (n < 2 || n > (1LL << 21)) return -1;
    const long long M = 1LL << bluestein_logm(n);
    return (int64_t)((n + 3 * M) * 8 + 256);
}

int ct_welch_single_f32(const float* x, int64_t n, double mean, int32_t use_abs, void* workspace, int64_t workspace_bytes,
                        double* acc, void* stream) {
    if (!x || !workspace || !acc) { ct_set_error("welch_single: null pointer"); return CT_ERR_ARG; }
    if (n < 2 || n > (1LL << 21)) { ct_set_error("welch_single: segment length must be in [2, 2^21] (got %lld)", (long long)n); return CT_ERR_UNSUPPORTED; }
    if (workspace_bytes < ct_welch_single_workspace_bytes(n)) { ct_set_error("welch_single: workspace too small"); return CT_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    const int m = bluestein_logm(n);
    const long long M = 1LL << m;
    const int logn1 = m / 2 < 10 ? m / 2 : 10, logn2 = m - logn1;
    cpx* c = (cpx*)workspace;
    cpx* bufA = (cpx*)((char*)workspace + (((size_t)n * 8 + 255) / 256) * 256);
    cpx* bufB = bufA + M;
    cpx* Y = bufB + M;
    const int T = 256;
    CT_COUNT_LAUNCH();
    ct_bluestein_chirp<<<(unsigned)((n + T - 1) / T), T, 0, st>>>(c, n);
    CT_COUNT_LAUNCH();
    ct_bluestein_kernel<<<(unsigned)((M + T - 1) / T), T, 0, st>>>(bufA, c, n, M);
    int rc = ct_check_launch("ct_bluestein_kernel"); if (rc) return rc;
    CfftArgs a;
    a.logn1 = logn1; a.logn2 = logn2; a.Y = Y;
    a.in = bufA; a.mul = nullptr; a.n_valid = M; a.conj_in = 0; a.out = bufB; a.conj_out = 0; a.scale = 1.f;
    rc = cfft(a, st); if (rc) return rc;                       // bufB = FFT(kernel)
    CT_COUNT_LAUNCH();
    ct_bluestein_input<<<(unsigned)((n + T - 1) / T), T, 0, st>>>(bufA, x, c, n, mean, use_abs);
    rc = ct_check_launch("ct_bluestein_input"); if (rc) return rc;
    a.in = bufA; a.n_valid = n; a.out = bufA;
    rc = cfft(a, st); if (rc) return rc;                       // bufA = FFT(x w c)
    a.in = bufA; a.mul = bufB; a.n_valid = M; a.conj_in = 1; a.out = bufA; a.conj_out = 1; a.scale = 1.0f / (float)M;
    rc = cfft(a, st); if (rc) return rc;                       // bufA = IFFT(A B) = conj(FFT(conj(A B))) / M
    const long long nout = n / 2 + 1;
    CT_COUNT_LAUNCH();
    ct_bluestein_power<<<(unsigned)((nout + T - 1) / T), T, 0, st>>>(acc, bufA, nout);
    return ct_check_launch("ct_bluestein_power");
}


}  // extern "C"
