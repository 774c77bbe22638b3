// draws.cuh -- device primitives of the draw specification v4.
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

namespace sg {

#define SG_HD __host__ __device__ __forceinline__
#define SG_D __device__ __forceinline__

constexpr double kTiny = 1e-160;              // pow(10,-160) clamps, src/gibbs_2.cpp:40-41,150,166,182,223
constexpr double kBinvMax = 30.0;             // n*p below this: inversion (as R's rbinom); otherwise BTRS
constexpr uint32_t kAuxBlock = 0xFFFFFFFFu;   // block holding (boost uniform, MH uniform)
constexpr uint32_t kBtrsBlock0 = 0x1000u;     // first BTRS block of a step-1 stream
constexpr int kSeg = 128;                     // columns per corrE segment

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
// log: fdlibm-style reduction to sqrt(1/2) < m < sqrt(2), degree-7 minimax in s = f/(2+f).
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
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    const int i = (hx + 0x95f64) & 0x100000;
    x = __hiloint2double(hx | (i ^ 0x3ff00000), (int)lx);
    k += (i >> 20);
    const double f = x - 1.0;
    const double s = f / (2.0 + f);
    const double dk = (double)k;
    const double z = s * s, w = z * z;
    const double t1 = w * fma(w, fma(w, 1.531383769920937332e-01, 2.222219843214978396e-01), 3.999999999940941908e-01);
    const double t2 = z * fma(w, fma(w, fma(w, 1.479819860511658591e-01, 1.818357216161805012e-01), 2.857142874366239149e-01),
                              6.666666666666735130e-01);
    const double R = t2 + t1;
    const double hfsq = 0.5 * f * f;
    return dk * 6.93147180369123816490e-01 - ((hfsq - (s * (hfsq + R) + dk * 1.90821492927058770002e-10)) - f);
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

// lgamma for x > 0: recurrence up to y >= 8, then 8 Stirling terms.
__device__ __noinline__ double det_lgamma(double x)
{
    if (x != x) return x;
    if (x < 0.0) return __longlong_as_double(0x7ff8000000000000LL);
    if (x == 0.0 || x > 1.7e308) return __longlong_as_double(0x7ff0000000000000LL);
    double y = x, corr = 0.0;
    if (x < 8.0) {
        double prod = x;
        prod = prod * (x + 1.0); prod = prod * (x + 2.0); prod = prod * (x + 3.0);
        prod = prod * (x + 4.0); prod = prod * (x + 5.0); prod = prod * (x + 6.0);
        prod = prod * (x + 7.0);
        corr = det_log(prod);
        y = x + 8.0;
    }
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

SG_D double horner8(double r, double c7, double c6, double c5, double c4, double c3, double c2, double c1, double c0)
{
    double v = fma(r, c7, c6);
    v = fma(v, r, c5); v = fma(v, r, c4); v = fma(v, r, c3);
    v = fma(v, r, c2); v = fma(v, r, c1); v = fma(v, r, c0);
    return v;
}

// Standard normal quantile, Wichura AS 241 PPND16.
SG_D double det_ppnd16(double p)
{
    const double q = p - 0.5;
    if (fabs(q) <= 0.425) {
        const double r = 0.180625 - q * q;
        const double num = horner8(r, 2509.0809287301226727, 33430.575583588128105, 67265.770927008700853, 45921.953931549871457,
                                   13731.693765509461125, 1971.5909503065514427, 133.14166789178437745, 3.387132872796366608);
        const double den = horner8(r, 5226.495278852545925, 28729.085735721942674, 39307.89580009271061, 21213.794301586595867,
                                   5394.1960214247511077, 687.1870074920579083, 42.313330701600911252, 1.0);
        return q * num / den;
    }
    double r = (q < 0) ? p : 1.0 - p;
    r = sqrt(-det_log(r));
    double val;
    if (r <= 5.0) {
        r -= 1.6;
        const double num = horner8(r, 7.7454501427834140764e-4, 0.0227238449892691845833, 0.24178072517745061177, 1.27045825245236838258,
                                   3.64784832476320460504, 5.7694972214606914055, 4.6303378461565452959, 1.42343711074968357734);
        const double den = horner8(r, 1.05075007164441684324e-9, 5.475938084995344946e-4, 0.0151986665636164571966, 0.14810397642748007459,
                                   0.68976733498510000455, 1.6763848301838038494, 2.05319162663775882187, 1.0);
        val = num / den;
    } else {
        r -= 5.0;
        const double num = horner8(r, 2.01033439929228813265e-7, 2.71155556874348757815e-5, 0.0012426609473880784386, 0.026532189526576123093,
                                   0.29656057182850489123, 1.7848265399172913358, 5.4637849111641143699, 6.6579046435011037772);
        const double den = horner8(r, 2.04426310338993978564e-15, 1.4215117583164458887e-7, 1.8463183175100546818e-5, 7.868691311456132591e-4,
                                   0.0148753612908506148525, 0.13692988092273580531, 0.59983220655588793769, 1.0);
        val = num / den;
    }
    return (q < 0) ? -val : val;
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

// ---------------------------------------------------------------- gamma (Marsaglia & Tsang 2000)
// Gamma(shape, rate) with one_gamma_dist's clamps (src/gibbs_2.cpp:39-43).  Attempt t reads block t.
__device__ __noinline__ double gamma_draw(const UStream &s, double shape, double rate)
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
        double u1, u2;
        s.pair(t, u1, u2);
        const double x = det_ppnd16(u1);
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
