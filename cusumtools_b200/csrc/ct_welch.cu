// Batched Welch periodogram with a hand-written shared-memory FFT (no cuFFT).
//
// Replaces scipy.signal.welch(x, fs, nperseg=L) as the reference calls it
// (plot-trace.py:442, noise-fit.py:92, legacy/minimal_psd.py:255): periodic Hann window,
// hop L/2, tail dropped, per-segment mean removal, density scaling, one-sided spectrum
// (scipy/signal/_spectral_py.py:515 -> csd -> ShortTimeFFT).  L is a power of two.
//
// A real segment of L samples is packed as N = L/2 complex points (z[j] = x[2j] + i x[2j+1])
// and transformed by the four-step algorithm, N = N1 x N2:
//   kernel A  window + pack + N2 column FFTs of length N1 (16 columns per CTA, data never
//             leaves shared memory) + twiddle W_N^(n2 k1) -> Y[k1][n2]      (4 B/sample in)
//   kernel B  row FFTs of length N2 for the row pair (k1, N1-k1), the real-FFT split
//             X[k] = (Z[k]+Z*[N-k])/2 - (i/2) e^(-i pi k/N) (Z[k]-Z*[N-k]) inside shared
//             memory, |X|^2 accumulated over the segments of the batch in registers and
//             added once per bin into the float64 accumulator.
// The intermediate Y (8 B per complex point) of a batch of segments is sized to stay in
// the 126 MB L2.  Mean removal is applied in the spectrum: with the input shifted by a
// constant c near the mean (for float32 headroom), FFT(w (x - mu)) = FFT(w (x - c)) -
// (mu - c) FFT(w) and FFT(w) of the periodic Hann is L/2 at bin 0 and -L/4 at bins +-1.
// FFT passes are Stockham autosort radix-8/4/2.  Every trigonometric factor (window, Stockham and
// four-step twiddles, real-FFT split) comes from ONE table T[j] = e^{-2 pi i j / L}, j < L, built once per
// call with sincospif (8 MB for L = 2^20, L2 resident); the sub-FFT twiddles are staged in shared memory.
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
template <> __device__ __forceinline__ void dft<2>(cpx* v) { cpx a = v[0]; v[0] = cadd(a, v[1]); v[1] = csub(a, v[1]); }
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

// One Stockham radix-R pass over `batch` interleaved FFTs of length N held as
// [index][batch] (batch fastest): in -> out.  Ns = product of the radices already done.
// tws[m] = e^{-2 pi i m / N} (shared memory): the pass twiddle e^{-2 pi i k r / (Ns R)} is tws[k r N/(Ns R)].
template <int R>
__device__ __forceinline__ void stockham_pass(const cpx* __restrict__ in, cpx* __restrict__ out, const cpx* __restrict__ tws,
                                              int N, int Ns, int batch, int tid, int nthreads) {
    const int work = (N / R) * batch;
    const int tstride = N / (Ns * R);
    for (int w = tid; w < work; w += nthreads) {
        const int b = w % batch, j = w / batch;
        const int k = j % Ns;
        cpx v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            cpx x = in[(j + r * (N / R)) * batch + b];
            v[r] = (r == 0 || Ns == 1) ? x : cmul(x, tws[k * r * tstride]);
        }
        dft<R>(v);
        const int j0 = (j / Ns) * Ns * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) out[(j0 + r * Ns) * batch + b] = v[r];
    }
}

// full FFT of length N = 2^logn for `batch` interleaved transforms; result pointer returned
__device__ cpx* fft_smem(cpx* a, cpx* b, const cpx* tws, int logn, int batch, int tid, int nthreads) {
    const int N = 1 << logn;
    int Ns = 1, rem = logn;
    while (rem > 0) {
        if (rem >= 3 && rem != 4) { stockham_pass<8>(a, b, tws, N, Ns, batch, tid, nthreads); Ns *= 8; rem -= 3; }
        else if (rem >= 2) { stockham_pass<4>(a, b, tws, N, Ns, batch, tid, nthreads); Ns *= 4; rem -= 2; }
        else { stockham_pass<2>(a, b, tws, N, Ns, batch, tid, nthreads); Ns *= 2; rem -= 1; }
        __syncthreads();
        cpx* t = a; a = b; b = t;
    }
    return a;
}

#ifndef CT_WELCH_COLS
#define CT_WELCH_COLS 8
#endif
constexpr int kCols = CT_WELCH_COLS;      // columns per CTA in kernel A (smem: 2 x N1 x kCols complex)
constexpr int kThreads = 256;

struct WelchArgs {
    const float* x; long long n;
    int L, logn1, logn2;       // N = L/2 = 2^logn1 * 2^logn2
    long long seg0; int nseg;  // segments [seg0, seg0+nseg) in this launch
    float c; int use_abs;
    const cpx* T;              // T[j] = e^{-2 pi i j / L}, j < L
    int rsplit;                // the segments of a batch are split over this many CTAs per row pair
    cpx* Y;                    // [nseg][N1][N2]
    double* segsum;            // [nseg] sum of (x - c) over the segment
    double* acc;               // [N+1]
    double mu_scale;           // 1/L
};

__global__ void ct_welch_table(cpx* T, int L) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < L) T[j] = expmi(2.0f * (float)j / (float)L);
}

__global__ void __launch_bounds__(kThreads) ct_welch_cols(WelchArgs a) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int N1 = 1 << a.logn1, N2 = 1 << a.logn2;
    cpx* A = reinterpret_cast<cpx*>(smraw);
    cpx* B = A + (size_t)N1 * kCols;
    cpx* tws = B + (size_t)N1 * kCols;             // e^{-2 pi i m / N1} = T[m * L / N1]
    for (int m = threadIdx.x; m < N1; m += kThreads) tws[m] = a.T[(size_t)m * (a.L / N1)];
    const int groups = N2 / kCols;
    const int s = blockIdx.x / groups, g = blockIdx.x % groups;
    const long long base = (a.seg0 + s) * (long long)(a.L / 2);   // hop = L/2
    const int tid = threadIdx.x;
    double part = 0.0;
    // load + window + pack: element (n1, col) is z[N2*n1 + g*16 + col] = x[2j], x[2j+1]
    for (int w = tid; w < N1 * kCols; w += kThreads) {
        const int col = w % kCols, n1 = w / kCols;
        const int j = N2 * n1 + g * kCols + col;
        float2 v = *reinterpret_cast<const float2*>(a.x + base + 2 * (long long)j);
        if (a.use_abs) { v.x = fabsf(v.x); v.y = fabsf(v.y); }
        v.x -= a.c; v.y -= a.c;
        part += (double)v.x + (double)v.y;
        const float4 t01 = *reinterpret_cast<const float4*>(a.T + 2 * j);      // T[2j], T[2j+1]: cos = .x, .z
        A[w] = make_float2(v.x * (0.5f - 0.5f * t01.x), v.y * (0.5f - 0.5f * t01.z));
    }
    // segment sum for the mean (warp + block reduction, one atomic per CTA)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(CT_FULL, part, o);
    __shared__ double wsum[kThreads / 32];
    if ((tid & 31) == 0) wsum[tid >> 5] = part;
    __syncthreads();
    if (tid == 0) {
        double t = 0; for (int i = 0; i < kThreads / 32; ++i) t += wsum[i];
        atomicAdd(a.segsum + s, t);
    }
    cpx* R = fft_smem(A, B, tws, a.logn1, kCols, tid, kThreads);
    // twiddle W_N^(n2 k1) and store Y[k1][n2]
    cpx* Y = a.Y + (size_t)s * N1 * N2;
    for (int w = tid; w < N1 * kCols; w += kThreads) {
        const int col = w % kCols, k1 = w / kCols;
        const int n2 = g * kCols + col;
        // W_N^(n2 k1) = T[2 (n2 k1 mod N)]   (n2 k1 < N1 N2 = N always)
        Y[(size_t)k1 * N2 + n2] = cmul(R[w], a.T[2 * (size_t)n2 * k1]);
    }
}

__global__ void __launch_bounds__(kThreads) ct_welch_rows(WelchArgs a) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int N1 = 1 << a.logn1, N2 = 1 << a.logn2;
    const long long N = (long long)N1 * N2;
    cpx* A = reinterpret_cast<cpx*>(smraw);       // [N2][2] interleaved pair of rows
    cpx* B = A + (size_t)N2 * 2;
    cpx* tws = B + (size_t)N2 * 2;                 // e^{-2 pi i m / N2} = T[m * L / N2]
    for (int m = threadIdx.x; m < N2; m += kThreads) tws[m] = a.T[(size_t)m * (a.L / N2)];
    const int r = blockIdx.x / a.rsplit;           // 0 .. N1/2
    const int part = blockIdx.x % a.rsplit;
    const int r2 = (r == 0) ? 0 : N1 - r;          // partner row (== r for r = 0 and N1/2)
    const bool self = (r2 == r);
    const int tid = threadIdx.x;
    // each thread owns bins k2 = tid, tid+256, ... of the row pair across all segments
    constexpr int kMaxOwn = 16;                    // N2 <= 4096
    float accA[kMaxOwn], accB[kMaxOwn];
#pragma unroll
    for (int i = 0; i < kMaxOwn; ++i) { accA[i] = 0.f; accB[i] = 0.f; }
    for (int s = part; s < a.nseg; s += a.rsplit) {
        const cpx* Y = a.Y + (size_t)s * N1 * N2;
        for (int w = tid; w < N2; w += kThreads) {
            A[2 * w] = Y[(size_t)r * N2 + w];
            A[2 * w + 1] = Y[(size_t)r2 * N2 + w];
        }
        __syncthreads();
        cpx* Z = fft_smem(A, B, tws, a.logn2, 2, tid, kThreads);
        const float dmu = (float)(a.segsum[s] * a.mu_scale);     // mu - c for this segment
        int own = 0;
        for (int k2 = tid; k2 < N2; k2 += kThreads, ++own) {
            // bin k = r + N1*k2 from row r, its mirror N-k from the partner row
            const long long k = r + (long long)N1 * k2;
            int m2; cpx Zk = Z[2 * k2], Zm;
            if (r == 0) { m2 = (N2 - k2) % N2; Zm = Z[2 * m2]; }
            else { m2 = N2 - 1 - k2; Zm = Z[2 * m2 + 1]; }
            cpx E = cadd(Zk, cconj(Zm)), O = csub(Zk, cconj(Zm));
            const cpx tw = a.T[k];                                // e^{-i pi k / N} = e^{-2 pi i k / L}
            cpx X = cadd(make_float2(0.5f * E.x, 0.5f * E.y), cmul(make_float2(0.5f * O.y, -0.5f * O.x), tw));
            // (-i/2) O tw  ==  (0.5*O.y, -0.5*O.x) * tw
            if (k == 0) X.x -= dmu * 0.5f * (float)a.L;
            if (k == 1) X.x += dmu * 0.25f * (float)a.L;
            accA[own] += X.x * X.x + X.y * X.y;
            // mirror bin N-k (k != 0): swap roles
            cpx E2 = cadd(Zm, cconj(Zk)), O2 = csub(Zm, cconj(Zk));
            const cpx tw2 = make_float2(-tw.x, tw.y);             // e^{-i pi (N-k)/N} = -conj(e^{-i pi k/N})
            cpx X2 = cadd(make_float2(0.5f * E2.x, 0.5f * E2.y), cmul(make_float2(0.5f * O2.y, -0.5f * O2.x), tw2));
            if (k == 0) {                                         // bin N (Nyquist) lives here
                X2 = make_float2(Zk.x - Zk.y, 0.f);
            }
            if (N - k == 1) X2.x += dmu * 0.25f * (float)a.L;
            accB[own] += X2.x * X2.x + X2.y * X2.y;
        }
        __syncthreads();
    }
    int own = 0;
    for (int k2 = tid; k2 < N2; k2 += kThreads, ++own) {
        const long long k = r + (long long)N1 * k2;
        const long long km = N - k;
        // every bin 0..N must be added exactly once per segment: row r owns k; the mirror
        // N-k is added here only if it belongs to the partner row of a non-self pair, or
        // (self-paired rows) if it is the Nyquist bin
        atomicAdd(a.acc + k, (double)accA[own]);
        if (!self) atomicAdd(a.acc + km, (double)accB[own]);
        else if (k == 0) atomicAdd(a.acc + N, (double)accB[own]);
    }
}

}  // namespace

extern "C" {

int64_t ct_welch_workspace_bytes(int32_t nperseg, int32_t batch) {
    if (nperseg < 256 || (nperseg & (nperseg - 1))) return -1;
    return (int64_t)batch * (nperseg / 2) * 8 + (int64_t)batch * 8 + 256 + (int64_t)nperseg * 8;
}

int ct_welch_f32(const float* x, int64_t n, int32_t nperseg, float shift, int32_t use_abs, int32_t batch,
                 void* workspace, int64_t workspace_bytes, double* acc, int64_t* nseg_out, void* stream) {
    if (!x || !workspace || !acc || !nseg_out) { ct_set_error("welch: null pointer"); return CT_ERR_ARG; }
    if (nperseg < 256 || (nperseg & (nperseg - 1)) || nperseg > (1 << 24)) {
        ct_set_error("welch: nperseg must be a power of two in [256, 2^24] (got %d)", nperseg); return CT_ERR_UNSUPPORTED;
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
    int logn2 = (logn + 1) / 2, logn1 = logn - logn2;      // N2 >= N1
    if (logn1 < 3) { logn1 = 3; logn2 = logn - 3; }
    if (logn2 > 12 || logn1 > 12 || (1 << logn2) < kCols) { ct_set_error("welch: unsupported segment length"); return CT_ERR_UNSUPPORTED; }
    WelchArgs a;
    a.x = x; a.n = n; a.L = L; a.logn1 = logn1; a.logn2 = logn2; a.c = shift; a.use_abs = use_abs;
    a.Y = (cpx*)workspace;
    a.segsum = (double*)((char*)workspace + (size_t)batch * N * 8);
    cpx* T = (cpx*)((char*)workspace + (((size_t)batch * N * 8 + (size_t)batch * 8 + 255) / 256) * 256);
    a.T = T;
    a.acc = acc; a.mu_scale = 1.0 / (double)L;
    const int N1 = 1 << logn1, N2 = 1 << logn2;
    CT_COUNT_LAUNCH();
    ct_welch_table<<<(L + 255) / 256, 256, 0, st>>>(T, L);
    { int rc = ct_check_launch("ct_welch_table"); if (rc) return rc; }
    size_t smA = (size_t)(2 * N1 * kCols + N1) * sizeof(cpx), smB = (size_t)(2 * 2 * N2 + N2) * sizeof(cpx);
    cudaFuncSetAttribute(ct_welch_cols, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smA);
    cudaFuncSetAttribute(ct_welch_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smB);
    for (long long s0 = 0; s0 < nseg; s0 += batch) {
        a.seg0 = s0; a.nseg = (int)((nseg - s0 < batch) ? nseg - s0 : batch);
        cudaMemsetAsync(a.segsum, 0, (size_t)a.nseg * 8, st);
        CT_COUNT_LAUNCH();
        ct_welch_cols<<<(unsigned)(a.nseg * (N2 / kCols)), kThreads, smA, st>>>(a);
        int rc = ct_check_launch("ct_welch_cols"); if (rc) return rc;
        // enough CTAs to fill the GPU: the segments of the batch are split over rsplit CTAs per row pair
        a.rsplit = 1;
        while ((N1 / 2 + 1) * a.rsplit < 4 * ct_sm_count() && a.rsplit * 2 <= a.nseg) a.rsplit *= 2;
        CT_COUNT_LAUNCH();
        ct_welch_rows<<<(unsigned)((N1 / 2 + 1) * a.rsplit), kThreads, smB, st>>>(a);
        rc = ct_check_launch("ct_welch_rows"); if (rc) return rc;
    }
    return CT_OK;
}

}  // extern "C"
