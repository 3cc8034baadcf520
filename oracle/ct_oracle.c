/* TEST INFRASTRUCTURE — plain-C twin of oracle/events_oracle.py and of scipy's DF2T lfilter.
 *
 * Same definitions, same order of individually rounded IEEE operations (compile with
 * -ffp-contract=off; x86-64 SSE2 arithmetic has no excess precision), so every output is
 * bit-identical to the NumPy definition while running ~1000x faster; tests assert that
 * identity and then use this twin at sizes NumPy loops cannot reach.  Never linked into or
 * loaded by the product library.
 *
 * Follows: scipy/signal/_signaltools.py:2268-2271 (_linear_filter, DF2T) for orc_lfilter;
 * plot-trace.py:379-414 (threshold / hysteresis lines) for orc_detect; SURVEY.md
 * Appendix C (two-sided CUSUM recurrence) for orc_cusum_batch.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* y = lfilter(b, a, x, zi=z); z updated in place.  ntaps = len(a) = len(b), a[0] == 1. */
void orc_lfilter(const double *b, const double *a, int ntaps, const double *x, int64_t n,
                 double *z, double *y)
{
    int m = ntaps - 1;
    for (int64_t i = 0; i < n; ++i) {
        double xi = x[i];
        double yi = z[0] + b[0] * xi;
        for (int j = 0; j < m - 1; ++j)
            z[j] = z[j + 1] + xi * b[j + 1] - yi * a[j + 1];
        z[m - 1] = xi * b[m] - yi * a[m];
        y[i] = yi;
    }
}

/* oracle/events_oracle.py::block_stats: a sample counts when its whole aligned 64-sample chunk (the last chunk of
 * the trace may be shorter) lies inside [bmin, bmax]; block % 64 == 0. */
void orc_block_stats(const float *y, int64_t n, int64_t block, float bmin, float bmax,
                     float c0, int shift, int64_t *cnt, int64_t *s1, int64_t *s2)
{
    float scale = ldexpf(1.0f, shift);
    int64_t nb = (n + block - 1) / block;
    for (int64_t k = 0; k < nb; ++k) {
        int64_t c = 0, a = 0, b = 0;
        int64_t e = (k + 1) * block < n ? (k + 1) * block : n;
        for (int64_t i0 = k * block; i0 < e; i0 += 64) {
            int64_t i1 = i0 + 64 < e ? i0 + 64 : e;
            int ok = 1;
            for (int64_t i = i0; i < i1; ++i)
                if (!(y[i] >= bmin && y[i] <= bmax)) { ok = 0; break; }
            if (!ok) continue;
            for (int64_t i = i0; i < i1; ++i) {
                float d = (y[i] - c0) * scale;
                int64_t q = (int64_t)rintf(d);
                c += 1; a += q; b += q * q;
            }
        }
        cnt[k] = c; s1[k] = a; s2[k] = b;
    }
}

/* returns number of complete events; *open_start = start of an event still open (or -1) */
int64_t orc_detect(const float *y, int64_t n, int64_t block, const int32_t *sign,
                   const float *t_start, const float *t_end, int state_in,
                   int64_t *starts, int64_t *ends, int64_t cap, int64_t *open_start)
{
    int inside = state_in != 0;
    int64_t start = -1, ne = 0;
    for (int64_t i = 0; i < n; ++i) {
        int64_t k = i / block;
        float v = y[i];
        int a, b;
        if (sign[k] > 0) { a = v < t_start[k]; b = v > t_end[k]; }
        else             { a = v > t_start[k]; b = v < t_end[k]; }
        if (!inside && a) { inside = 1; start = i; }
        else if (inside && b) {
            inside = 0;
            if (start >= 0) {
                if (ne < cap) { starts[ne] = start; ends[ne] = i; }
                ++ne;
            }
        }
    }
    *open_start = inside ? start : -1;
    return ne;
}

#define CQ 64.0f
#define CSSCALE 1024.0f
#define CSMAX 2097152.0f

static inline int64_t quantise(float x, float x0)
{
    float d = (x - x0) * CQ;
    d = rintf(d);
    if (!(d > -2147483648.0f)) return d != d ? 0 : -2147483648LL;     /* the conversion saturates at the int32 range */
    if (d >= 2147483648.0f) return 2147483647LL;
    return (int64_t)d;
}

/* One event.  edges[0..max_levels], returns number of levels; *overflow set if truncated. */
static int cusum_event(const float *x, int64_t n, float delta, float h, int max_levels,
                       int32_t *edges, double *mean, double *std, uint8_t *overflow)
{
    int64_t H = (int64_t)rintf(h * CSSCALE);
    float dq = delta * CQ;
    float hq = dq * 0.5f;
    float x0 = x[0];
    int nedge = 0;
    edges[nedge++] = 0;
    *overflow = 0;
    int64_t k0 = 0;
    int64_t qa = quantise(x[0], x0);
    uint64_t Sd = 0, Sdd = 0;                        /* two's-complement sums of d and d^2 (d = q - q_anchor) */
    int64_t gp = 0, gn = 0, rp = 0, rn = 0;
    for (int64_t k = 1; k < n; ++k) {
        int64_t d = quantise(x[k], x0) - qa;
        Sd += (uint64_t)d; Sdd += (uint64_t)d * (uint64_t)d;
        int64_t cnt = k - k0 + 1;
        float rc = 1.0f / (float)cnt;
        float m = (float)(int64_t)Sd * rc;
        float p1 = (float)(int64_t)Sdd * rc;
        float p2 = m * m;
        float v = p1 - p2;
        int64_t sp = 0, sn = 0;
        if (v > 0.0f) {
            float ri = 1.0f / v;
            float r = dq * ri;
            float t = (float)d - m;
            float a = (r * (t - hq)) * CSSCALE;
            float b = ((-r) * (t + hq)) * CSSCALE;
            a = fminf(fmaxf(a, -CSMAX), CSMAX);
            b = fminf(fmaxf(b, -CSMAX), CSMAX);
            sp = (int64_t)rintf(a);
            sn = (int64_t)rintf(b);
        }
        gp += sp; if (gp <= 0) { gp = 0; rp = k; }
        gn += sn; if (gn <= 0) { gn = 0; rn = k; }
        if (gp > H || gn > H) {
            int64_t jmin = gp >= gn ? rp : rn;
            if (nedge >= max_levels) { *overflow = 1; break; }
            edges[nedge++] = (int32_t)(jmin + 1);
            k0 = k; qa += d; Sd = 0; Sdd = 0; gp = gn = 0; rp = rn = k;
        }
    }
    edges[nedge++] = (int32_t)n;
    int L = nedge - 1;
    for (int i = 0; i < L; ++i) {
        int64_t a = 0, b = 0;
        for (int64_t j = edges[i]; j < edges[i + 1]; ++j) {
            int64_t qj = quantise(x[j], x0);
            a += qj; b += qj * qj;
        }
        double len = (double)(edges[i + 1] - edges[i]);
        double ad = (double)a, bd = (double)b;
        mean[i] = (double)x0 + (ad / len) / 64.0;
        double var = (bd - ad * ad / len);
        if (var < 0.0) var = 0.0;
        std[i] = sqrt(var / len) / 64.0;
    }
    return L;
}

void orc_cusum_batch(const float *samples, const int64_t *offsets, int64_t nevents,
                     float delta, float h, int max_levels, int32_t *n_levels,
                     int32_t *edges /*[E][max_levels+1]*/, double *mean /*[E][max_levels]*/,
                     double *std, uint8_t *overflow)
{
    for (int64_t e = 0; e < nevents; ++e) {
        int32_t *ed = edges + e * (max_levels + 1);
        for (int i = 0; i <= max_levels; ++i) ed[i] = -1;
        int64_t n = offsets[e + 1] - offsets[e];
        n_levels[e] = 0; overflow[e] = 0;
        if (n <= 0) continue;
        n_levels[e] = cusum_event(samples + offsets[e], n, delta, h, max_levels, ed,
                                  mean + e * max_levels, std + e * max_levels, overflow + e);
    }
}
