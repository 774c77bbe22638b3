// draws.cuh -- device primitives of the draw specification v7.
//
// Everything here is built from IEEE-754 correctly rounded fp64 operations (+ - * / sqrt fma) in
// a fixed order, plus integer arithmetic, so that results are bit-identical to any other
// conforming implementation (tests compare against the CPU oracle).  This translation unit is
// compiled with --fmad=false: the only fused multiply-adds are the explicit fma() calls.
//
// Replaces the third-party arithmetic the reference calls (not in its repo): R nmath
// rmultinom/rbinom/rgamma/unif_rand behind src/gibbs_2.cpp:34,42 and arma::lgamma/log/exp/randu
// behind src/gibbs_2.cpp:186-193,227-234.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "det_tables.h"

namespace sg {

#define SG_HD __host__ __device__ __forceinline__
#define SG_D __device__ __forceinline__

constexpr double kTiny = 1e-160;              // pow(10,-160) clamps, src/gibbs_2.cpp:40-41,150,166,182,223
#ifndef SG_BINV_MAX
#define SG_BINV_MAX 30.0
#endif
constexpr double kBinvMax = SG_BINV_MAX;      // n*p below this: inversion (as R's rbinom); otherwise BTRS
constexpr uint32_t kAuxBlock = 0xFFFFFFFFu;   // block holding (boost uniform, MH uniform)
constexpr uint32_t kBtrsBlock0 = 0x1000u;     // first BTRS block of a step-1 stream
constexpr int kSeg = 32;                      // genomes per W E' segment (draw specification v7)

enum Stream : uint32_t { kStPerm = 0, kStZ = 1, kStP = 2, kStE = 3, kStBp = 4, kStBe = 5, kStAp = 6, kStAe = 7, kStZinit = 8 };

// ---------------------------------------------------------------- Philox4x32-10 (Salmon et al. 2011)
struct U4 { uint32_t x, y, z, w; };

SG_HD U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), h1 = __umulhi(0xCD9E8D57u, c2);
#else
        const uint32_t h0 = (uint32_t)(((uint64_t)0xD2511F53u * c0) >> 32);
        const uint32_t h1 = (uint32_t)(((uint64_t)0xCD9E8D57u * c2) >> 32);
#endif
        const uint32_t l0 = 0xD2511F53u * c0, l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return U4{c0, c1, c2, c3};
}

// (2 * x52 + 1) * 2^-53 with x52 = top 26 bits of w0 : top 26 bits of w1, built without an
// integer->double conversion: 1.x52 - 1 is exact, adding 2^-53 is exact (53 significant bits).
SG_D double u53(uint32_t w0, uint32_t w1)
{
    const uint32_t hi = 0x3ff00000u | (w0 >> 12);
    const uint32_t lo = ((w0 >> 6) << 26) | (w1 >> 6);
    return (__hiloint2double((int)hi, (int)lo) - 1.0) + 0x1p-53;
}

// (w + 1/2) * 2^-32 in (0,1): 32 random bits, the resolution of R's unif_rand (Mersenne-Twister).
// 1.(w << 20) - 1 = w * 2^-32 exactly; adding 2^-33 is exact (33 significant bits).
SG_D double u32(uint32_t w)
{
    return (__hiloint2double((int)(0x3ff00000u | (w >> 12)), (int)(w << 20)) - 1.0) + 0x1p-33;
}

struct Key { uint32_t k0, k1; };

// A stream is addressed by (key, sweep, stream id, element); its blocks are addressed directly.
struct UStream {
    uint32_t k0, k1, elem, stream, sweep;
    SG_D void init(Key k, uint32_t sw, uint32_t st, uint32_t e)
    {
        k0 = k.k0; k1 = k.k1; elem = e; stream = st; sweep = sw;
    }
    SG_D void pair(uint32_t b, double &u0, double &u1) const
    {
        const U4 w = philox4x32_10(b, elem, stream, sweep, k0, k1);
        u0 = u53(w.x, w.y);
        u1 = u53(w.z, w.w);
    }
    SG_D U4 words(uint32_t b) const { return philox4x32_10(b, elem, stream, sweep, k0, k1); }
};

// ---------------------------------------------------------------- deterministic elementary functions
// log, division-free: x = 2^k z with bits(z) in [OFF, OFF + 2^52) (z in [0.6875, 1.375)); a 128-entry table gives
// invc ~ 1/z and logc = -log(invc) for the run of bit patterns z falls in; r = fma(z, invc, -1), |r| < 2^-7;
// log x = k ln2 + logc + log1p(r), log1p by its Taylor polynomial up to r^9, the leading terms summed in two pieces.
// Table: scripts/make_det_tables.py (det_tables.h; the oracle carries a byte-identical copy).
__device__ const double g_log_tab[SG_LOG_N][2] = SG_LOG_TABLE;

SG_D double det_log(double x)
{
    int hx = __double2hiint(x);
    unsigned lx = (unsigned)__double2loint(x);
    int k = 0;
    if (hx < 0x00100000) {
        if (((hx & 0x7fffffff) | (int)lx) == 0) return -__longlong_as_double(0x7ff0000000000000LL) * 1.0;
        if (hx < 0) return __longlong_as_double(0x7ff8000000000000LL);
        k = -54; x = x * 0x1p54; hx = __double2hiint(x); lx = (unsigned)__double2loint(x);
    }
    if (hx >= 0x7ff00000) return x + x;
    const int tmp = hx - (int)SG_LOG_OFF_HI;
    const int i = (tmp >> 13) & (SG_LOG_N - 1);
    k += tmp >> 20;
    const double z = __hiloint2double(hx - (tmp & (int)0xfff00000), (int)lx);
    const double2 t = __ldg(reinterpret_cast<const double2 *>(&g_log_tab[i][0]));    // {invc, logc}
    const double r = fma(z, t.x, -1.0);
    const double kd = (double)k;
    const double w = fma(kd, 6.93147180369123816490e-01, t.y);
    const double hi = w + r;
    const double lo = ((w - hi) + r) + kd * 1.90821492927058770002e-10;
    double q = fma(r, 1.0 / 9.0, -1.0 / 8.0);
    q = fma(r, q, 1.0 / 7.0); q = fma(r, q, -1.0 / 6.0); q = fma(r, q, 1.0 / 5.0);
    q = fma(r, q, -1.0 / 4.0); q = fma(r, q, 1.0 / 3.0); q = fma(r, q, -0.5);
    return (lo + (r * r) * q) + hi;
}

SG_D double det_exp(double x)
{
    if (x != x) return x;
    if (x > 7.09782712893383973096e+02) return __longlong_as_double(0x7ff0000000000000LL);
    if (x < -7.45133219101941108420e+02) return 0.0;
    double hi = x, lo = 0.0, xr = x;
    int k = 0;
    if (fabs(x) > 0.34657359027997264) {
        k = (int)(1.44269504088896338700e+00 * x + (x < 0 ? -0.5 : 0.5));
        const double t = (double)k;
        hi = x - t * 6.93147180369123816490e-01;
        lo = t * 1.90821492927058770002e-10;
        xr = hi - lo;
    }
    const double t = xr * xr;
    const double c = xr - t * fma(t, fma(t, fma(t, fma(t, 4.13813679705723846039e-08, -1.65339022054652515390e-06),
                                                6.61375632143793436117e-05), -2.77777777770155933842e-03), 1.66666666666666019037e-01);
    const double y = 1.0 - ((lo - (xr * c) / (2.0 - c)) - hi);
    if (k == 0) return y;
    if (k >= -1021) {
        if (k == 1024) return y * 2.0 * 0x1p1023;
        return __longlong_as_double(__double_as_longlong(y) + ((long long)k << 52));
    }
    return __longlong_as_double(__double_as_longlong(y) + ((long long)(k + 1000) << 52)) * 0x1p-1000;
}

// lgamma for x > 0: recurrence up to y >= 8, then 8 Stirling terms.  Written without a data-dependent branch (the
// recurrence runs on a dummy argument when x >= 8 and its logarithm is discarded) so that several evaluations inline
// into one straight-line block and overlap; the selected operations and their order are the specification's.
SG_D double det_lgamma(double x)
{
    if (x != x) return x;
    if (x < 0.0) return __longlong_as_double(0x7ff8000000000000LL);
    if (x == 0.0 || x > 1.7e308) return __longlong_as_double(0x7ff0000000000000LL);
    const bool small = x < 8.0;
    const double xs = small ? x : 1.0;
    double prod = xs;
    prod = prod * (xs + 1.0); prod = prod * (xs + 2.0); prod = prod * (xs + 3.0);
    prod = prod * (xs + 4.0); prod = prod * (xs + 5.0); prod = prod * (xs + 6.0);
    prod = prod * (xs + 7.0);
    const double lp = det_log(prod);
    const double corr = small ? lp : 0.0;
    const double y = small ? x + 8.0 : x;
    const double r = 1.0 / y, r2 = r * r;
    double ser = fma(r2, -2.9550653594771242e-02, 6.4102564102564100e-03);
    ser = fma(r2, ser, -1.9175269175269176e-03);
    ser = fma(r2, ser, 8.4175084175084171e-04);
    ser = fma(r2, ser, -5.9523809523809529e-04);
    ser = fma(r2, ser, 7.9365079365079365e-04);
    ser = fma(r2, ser, -2.7777777777777779e-03);
    ser = fma(r2, ser, 8.3333333333333329e-02);
    ser = r * ser;
    const double lg = ((y - 0.5) * det_log(y) - y) + 0.91893853320467274178 + ser;
    return lg - corr;
}

// ---------------------------------------------------------------- binomial
__constant__ double c_btrs_fc[10] = {
    0.08106146679532726, 0.04134069595540929, 0.02767792568499834, 0.02079067210376509,
    0.01664469118982119, 0.01387612882307075, 0.01189670994589177, 0.01041126526197209,
    0.009255462182712733, 0.008330563433362871 };

SG_D double btrs_fc(double k)
{
    if (k <= 9.0) return c_btrs_fc[(int)k];
    const double ri = 1.0 / (k + 1.0), ri2 = ri * ri;
    return (1.0 / 12.0 - (1.0 / 360.0 - (1.0 / 1260.0) * ri2) * ri2) * ri;
}

// 1/x for x = 1..255 (entry 0 unused): the inversion recurrence multiplies by the correctly
// rounded reciprocal instead of dividing.
struct RcpTab {
    double v[256];
    constexpr RcpTab() : v{} { for (int i = 1; i < 256; ++i) v[i] = 1.0 / (double)i; }
};
__device__ const RcpTab g_rcp_tab = RcpTab();

SG_D double rcp_int(int x, double xd) { return x < 256 ? __ldg(&g_rcp_tab.v[x]) : 1.0 / xd; }

// Hormann's BTRS (transformed rejection with squeeze) for n*p >= kBinvMax, p <= 1/2.  Out of line:
// it is rare and register-hungry.  Attempt t reads block kBtrsBlock0 + (kstep<<8) + t.
__device__ __noinline__ int btrs_draw(UStream s, int kstep, int n, double p)
{
    const double q = 1.0 - p, nd = (double)n;
    const double spq = sqrt(nd * p * q);
    const double b = 1.15 + 2.53 * spq;
    const double a = -0.0873 + 0.0248 * b + 0.01 * p;
    const double c = nd * p + 0.5;
    const double vr = 0.92 - 4.2 / b;
    const double alpha = (2.83 + 5.1 / b) * spq;
    const double r = p / q;
    const double m = floor((nd + 1.0) * p);
    double kd = 0.0;
    for (uint32_t t = 0; t < 256; ++t) {
        const U4 w = s.words(kBtrsBlock0 + ((uint32_t)kstep << 8) + t);
        const double U = u32(w.x) - 0.5, V = u32(w.y);
        const double us = 0.5 - fabs(U);
        kd = floor((2.0 * a / us + b) * U + c);
        if (us >= 0.07 && V <= vr) return (int)kd;
        if (kd < 0.0 || kd > nd) continue;
        const double lv = det_log(V * alpha / (a / (us * us) + b));
        const double ub = (m + 0.5) * det_log((m + 1.0) / (r * (nd - m + 1.0)))
                        + (nd + 1.0) * det_log((nd - m + 1.0) / (nd - kd + 1.0))
                        + (kd + 0.5) * det_log(r * (nd - kd + 1.0) / (kd + 1.0))
                        + btrs_fc(m) + btrs_fc(nd - m) - btrs_fc(kd) - btrs_fc(nd - kd);
        if (lv <= ub) return (int)kd;
    }
    return (kd < 0.0) ? 0 : (kd > nd) ? n : (int)kd;
}

// Binomial(n, pp) for n >= 1, 0 < pp < 1, conditional binomial number `kstep` of its cell.
// u_binv: the inversion uniform = u32(word (kstep&3) of block (kstep>>2)) of the cell's stream (the
// caller generates the block, warp-converged, once per four steps).
SG_D int binomial_draw(const UStream &s, int kstep, double u_binv, int n, double pp)
{
    const bool flip = pp > 0.5;
    const double p = flip ? 1.0 - pp : pp;
    const double q = 1.0 - p;
    const double nd = (double)n;
    int x = 0;
    if (nd * p < kBinvMax) {
        // inversion by sequential search from P(X=0) = q^n (binary powering)
        double r = 1.0, b = q;
        unsigned e = (unsigned)n;
        for (;;) {
            if (e & 1u) r = r * b;
            e >>= 1;
            if (!e) break;
            b = b * b;
        }
        double u = u_binv;
        if (u >= r) {
            const double sr = p / q;
            const double a = sr * (nd + 1.0);         // sr*(n-x+1)/x = a/x - sr
            double xd = 0.0;
            while (u >= r && x < n) {
                u = u - r;
                x += 1; xd += 1.0;
                r = r * fma(a, rcp_int(x, xd), -sr);
            }
        }
    } else {
        x = btrs_draw(s, kstep, n, p);
    }
    return flip ? n - x : x;
}

// ---------------------------------------------------------------- normal (ziggurat)
// Marsaglia & Tsang 2000 with the layer index independent of the value (Doornik 2005): 128 layers of equal area with
// right edges g_zig_x[0..128].  Candidate from four Philox words: |u| = the 52 bits (w1 >> 12 : w0) as a fraction,
// sign = bit 7 of w1, layer = the low 7 bits of w1; accepted at once when |u| < x[i+1]/x[i] (97 %).  Otherwise the
// base strip draws from the tail beyond R and the other layers test the wedge under the density.  Attempt t of a gamma
// draw takes its first candidate from block t; further uniforms / candidates take blocks kNormBlock0 + (t << 8) + q.
__device__ const double g_zig_x[129] = SG_ZIG_X;
__device__ const double g_zig_ratio[128] = SG_ZIG_RATIO;
constexpr uint32_t kNormBlock0 = 0x10000u;

__device__ __noinline__ double normal_slow(const UStream s, uint32_t t, U4 w)
{
    uint32_t q = 0;
    for (;;) {
        const double mag = __hiloint2double((int)(0x3ff00000u | (w.y >> 12)), (int)w.x) - 1.0;
        const int i = (int)(w.y & 127u);
        const bool neg = (w.y >> 7) & 1u;
        const double xi = __ldg(&g_zig_x[i]);
        if (mag < __ldg(&g_zig_ratio[i])) { const double x = mag * xi; return neg ? -x : x; }
        if (i == 0) {
            double xx, yy;
            do {
                double u1, u2;
                s.pair(kNormBlock0 + (t << 8) + q, u1, u2); ++q;
                xx = det_log(u1) / SG_ZIG_R;
                yy = det_log(u2);
            } while (-2.0 * yy < xx * xx);
            return neg ? xx - SG_ZIG_R : SG_ZIG_R - xx;
        }
        const double x = mag * xi, xi1 = __ldg(&g_zig_x[i + 1]);
        const double f0 = det_exp(-0.5 * (xi * xi - x * x)), f1 = det_exp(-0.5 * (xi1 * xi1 - x * x));
        double ua, ub;
        s.pair(kNormBlock0 + (t << 8) + q, ua, ub); ++q;
        if (f1 + ua * (f0 - f1) < 1.0) return neg ? -x : x;
        w = s.words(kNormBlock0 + (t << 8) + q); ++q;
    }
}

// w: the four words of block t
SG_D double normal_draw(const UStream &s, uint32_t t, const U4 &w)
{
    const double mag = __hiloint2double((int)(0x3ff00000u | (w.y >> 12)), (int)w.x) - 1.0;
    const int i = (int)(w.y & 127u);
    if (mag < __ldg(&g_zig_ratio[i])) {
        const double x = mag * __ldg(&g_zig_x[i]);
        return ((w.y >> 7) & 1u) ? -x : x;
    }
    return normal_slow(s, t, w);
}

// ---------------------------------------------------------------- gamma (Marsaglia & Tsang 2000)
// Gamma(shape, rate) with one_gamma_dist's clamps (src/gibbs_2.cpp:39-43).  Attempt t reads block t: the normal
// from words 0,1 (normal_draw), the acceptance uniform from words 2,3.
__device__ __noinline__ double gamma_draw(const UStream s, double shape, double rate)
{
    const double a = (kTiny < shape) ? shape : kTiny;
    const double invr = 1.0 / rate;
    const double scale = (kTiny < invr) ? invr : kTiny;
    const bool boost = a < 1.0;
    const double aa = boost ? a + 1.0 : a;
    const double d = aa - 1.0 / 3.0;
    const double c = 1.0 / sqrt(9.0 * d);
    double v = 1.0;
    for (uint32_t t = 0; t < 256; ++t) {
        const U4 w = s.words(t);
        const double u2 = u53(w.z, w.w);
        const double x = normal_draw(s, t, w);
        v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        const double x2 = x * x;
        if (u2 < 1.0 - 0.0331 * (x2 * x2)) break;
        if (det_log(u2) < 0.5 * x2 + d * (1.0 - v + det_log(v))) break;
    }
    double g = d * v;
    if (boost) {
        double ub, um;
        s.pair(kAuxBlock, ub, um);
        g = g * det_exp(det_log(ub) / a);
    }
    return g * scale;
}

} // namespace sg
