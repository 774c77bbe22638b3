/*
 * oracle/gibbs_oracle.c -- CPU oracle for signeR's empirical-Bayes Poisson-NMF Gibbs sampler.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (signer_b200/, include/) may link,
 * import or execute this file.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs use it, and only as the checker / the timed CPU baseline.
 *
 * PARITY UNPINNED: the reference (TojalLab/signeR v2.3.3) ships no tests, no golden vectors and
 * cannot be built here (needs R, Rcpp, RcppArmadillo, R nmath).  What this file restates:
 *
 *   - the update equations, clamps, special cases and storage rules of
 *       src/gibbs_2.cpp:85-103   (step 1, Z | P,E  -- sequential conditional binomials, as R's rmultinom)
 *       src/gibbs_2.cpp:106-137  (steps 2,3: conjugate Gamma draws of P and E)
 *       src/gibbs_2.cpp:140-169  (steps 4,5: Bp, Be, clamp at 1e-160)
 *       src/gibbs_2.cpp:172-251  (steps 6,7: Metropolis-Hastings on Ap, Ae)
 *       src/gibbs_2.cpp:39-43    (one_gamma_dist clamps)
 *       src/gibbs_2.cpp:303-318  (random scan: fresh permutation of steps 1..7 every sweep, guards)
 *       src/gibbs_2.cpp:320-334  (sample storage for sweeps >= burn; Zs keeps only the latest Z)
 *       src/gibbs_2.cpp:9-28     (cube sums, slice-innermost loops over a materialised I x J x K cube)
 *       R/cpp_eBayesNMF.R:180-186 (per-sample Poisson log-likelihood; BIC = 2 ll - n (i+j) log j)
 *   - R's rmultinom structure (nmath/rmultinom.c, not in the repo):
 *       k order, zero-probability skip, pp<1 test, early exit at n<=0, remainder to the last class.
 *
 * What it does NOT reproduce: R's RNG stream and R's rbinom/rgamma algorithms (unavailable).
 * Instead every random draw follows the "draw specification v7" below: counter-based
 * Philox4x32-10 streams addressed by (seed, sweep, step, element, block) and samplers built only
 * from IEEE-754 correctly rounded operations (+ - * / sqrt fma) in a fixed order, so that an
 * independent implementation (the CUDA kernels) can be compared BIT FOR BIT.
 * Compile with -ffp-contract=off (see oracle/Makefile); every fused multiply-add is explicit.
 *
 * Draw specification v7 (shared contract with signer_b200/csrc, written independently there):
 *   stream address: key = (seed_lo, seed_hi); counter = (block, element, stream_id, sweep)
 *   stream ids: 0 permutation, 1..7 Gibbs steps, 8 initial Z
 *   element: step 1/8: i + I*j ; P family (2,4,6): i + I*k ; E family (3,5,7): k + K*j
 *   u53(w0,w1) = (2*((w0>>6)<<26 | (w1>>6)) + 1) * 2^-53       in (0,1)
 *   u32(w) = (w + 1/2) * 2^-32: step-1 uniforms carry 32 random bits like R's unif_rand (Mersenne-Twister)
 *   step 1, cells with count n <= 32: direct method, n categorical draws; draw t uses word (t&3) of
 *                block (t>>2): x = u*S, class = first k < K-1 with x < c_k, c_k = c_{k-1} + P[i,k]*E[k,j],
 *                else the last class.
 *   step 1, larger cells: sequential conditional binomials (rmultinom's rules); the classes are visited in
 *                the order of d_k = min(7, expo(S) - expo(P*E)) ascending (expo = biased binary exponent),
 *                ties lowest index first; conditional binomial number t: the inversion uniform is word (t&3)
 *                of block (t>>2); BTRS attempt a of binomial t reads block 0x1000 + (t<<8) + a (U from word
 *                0, V from word 1).  The chain stops once at most 32 mutations are left; those are dealt by
 *                the direct method over the unvisited classes: c_k in index order with the visited classes'
 *                P*E set to zero, Sr = c_{K-1}, draw d uses word (d&3) of block 0x800 + (d>>2), x = u*Sr,
 *                class = first k < K-1 with x < c_k, else K-1.
 *   binomial(n,pp): p=min(pp,1-pp); n*p<30 -> inversion from q^n (binary powering), recurrence
 *                   r <- r * fma(a, 1/x, -s), s = p/q, a = s*(n+1); else Hormann BTRS; flip back if pp>0.5
 *   normal:      ziggurat, 128 layers (det_tables.h), see normal_draw
 *   gamma draws: Marsaglia-Tsang attempt t uses block t: normal from words 0,1 (further blocks
 *                0x10000 + (t<<8) + q when the fast path fails), accept test u53(w2,w3);
 *                block 0xFFFFFFFF gives (boost uniform for shape<1, MH acceptance uniform)
 *   log:         table-driven, division-free (det_tables.h), see oracle_log; exp: fdlibm-style; lgamma: shift to
 *                >= 8 + 8 Stirling terms
 *   fp reductions: PE = fma chain over k; corrP = fma chain over i; corrE = fma chains over segments of 32 genomes,
 *                  the segment sums added in blocks of 32 (sequentially inside a block; the block sums form the
 *                  next level) per shard, the shard sums likewise;
 *                  loglik = per-column sums in 32-row chunks (chunks added in order), then a
 *                  stride-halving tree over the zero-padded column array.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "det_tables.h"

#define ORACLE_SEG 32
#define ORACLE_SUMBLK 32
#define ORACLE_LLCHUNK 32
#define BINV_MAX 30.0
#define TINY 1e-160
#define AUX_BLOCK 0xFFFFFFFFu
#define BTRS_BLOCK0 0x1000u
#define NDIRECT 32
#define TAIL_BLOCK0 0x800u
#define ORACLE_KMAX 128

enum { ST_PERM = 0, ST_Z = 1, ST_P = 2, ST_E = 3, ST_BP = 4, ST_BE = 5, ST_AP = 6, ST_AE = 7, ST_ZINIT = 8 };

/* ------------------------------------------------------------------ Philox4x32-10 */
/* Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11); Random123 philox.h.
 * Known-answer vectors are checked in tests/test_oracle_kat.py. */
void oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

double oracle_u53(uint32_t w0, uint32_t w1)
{
    uint64_t x = ((uint64_t)(w0 >> 6) << 26) | (uint64_t)(w1 >> 6);
    return (double)(2 * x + 1) * 0x1p-53;
}

/* 32-bit uniform (w + 1/2) * 2^-32 in (0,1): the resolution of R's own unif_rand() under its
 * default generator (Mersenne-Twister, 32 random bits per uniform), used by every step-1 draw. */
double oracle_u32(uint32_t w) { return ((double)w + 0.5) * 0x1p-32; }

typedef struct {
    uint32_t key[2];
    uint32_t elem, stream, sweep;
} ustream;

static void us_init(ustream *s, uint64_t seed, uint32_t sweep, uint32_t stream, uint32_t elem)
{
    s->key[0] = (uint32_t)seed; s->key[1] = (uint32_t)(seed >> 32);
    s->elem = elem; s->stream = stream; s->sweep = sweep;
}

static void us_quad(const ustream *s, uint32_t block, double u[4])
{
    uint32_t ctr[4] = { block, s->elem, s->stream, s->sweep }, w[4];
    oracle_philox(ctr, s->key, w);
    for (int i = 0; i < 4; i++) u[i] = oracle_u32(w[i]);
}

static void us_block(const ustream *s, uint32_t block, double *u0, double *u1)
{
    uint32_t ctr[4] = { block, s->elem, s->stream, s->sweep }, w[4];
    oracle_philox(ctr, s->key, w);
    *u0 = oracle_u53(w[0], w[1]);
    *u1 = oracle_u53(w[2], w[3]);
}

/* ------------------------------------------------------------------ deterministic math */
static inline uint64_t d2u(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double u2d(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }

/* log(x), division-free: x = 2^k z with bits(z) in [OFF, OFF + 2^52) (z in [0.6875, 1.375)); a 128-entry table gives
 * invc ~ 1/z and logc = -log(invc) for the run of bit patterns z falls in; r = fma(z, invc, -1) (|r| < 2^-7);
 * log x = k ln2 + logc + log1p(r), log1p by its Taylor polynomial up to r^9, the leading terms summed in two pieces.
 * (The scheme of table-driven logarithms, e.g. Tang 1990; the table is generated by scripts/make_det_tables.py.) */
static const double log_tab[SG_LOG_N][2] = SG_LOG_TABLE;

double oracle_log(double x)
{
    static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    uint64_t ux = d2u(x);
    int32_t hx = (int32_t)(ux >> 32);
    uint32_t lx = (uint32_t)ux;
    int k = 0;
    if (hx < 0x00100000) {
        if (((hx & 0x7fffffff) | lx) == 0) return -INFINITY;
        if (hx < 0) return NAN;
        k -= 54; x *= 0x1p54; ux = d2u(x); hx = (int32_t)(ux >> 32); lx = (uint32_t)ux;
    }
    if (hx >= 0x7ff00000) return x + x;
    int32_t tmp = hx - (int32_t)SG_LOG_OFF_HI;
    int i = (tmp >> 13) & (SG_LOG_N - 1);
    k += tmp >> 20;                                        /* arithmetic shift */
    double z = u2d(((uint64_t)(uint32_t)(hx - (tmp & (int32_t)0xfff00000)) << 32) | lx);
    double invc = log_tab[i][0], logc = log_tab[i][1];
    double r = fma(z, invc, -1.0);
    double kd = (double)k;
    double w = fma(kd, ln2_hi, logc);
    double hi = w + r;
    double lo = ((w - hi) + r) + kd * ln2_lo;
    double q = fma(r, 1.0 / 9.0, -1.0 / 8.0);
    q = fma(r, q, 1.0 / 7.0); q = fma(r, q, -1.0 / 6.0); q = fma(r, q, 1.0 / 5.0);
    q = fma(r, q, -1.0 / 4.0); q = fma(r, q, 1.0 / 3.0); q = fma(r, q, -0.5);
    return (lo + (r * r) * q) + hi;
}

/* exp(x): fdlibm e_exp.c style reduction x = k ln2 + r, rational form, unified formula. */
double oracle_exp(double x)
{
    static const double ln2HI = 6.93147180369123816490e-01, ln2LO = 1.90821492927058770002e-10,
        invln2 = 1.44269504088896338700e+00,
        P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
        P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
    if (x != x) return x;
    if (x > 7.09782712893383973096e+02) return INFINITY;
    if (x < -7.45133219101941108420e+02) return 0.0;
    double hi, lo, xr;
    int k;
    if (fabs(x) > 0.34657359027997264) {
        k = (int)(invln2 * x + (x < 0 ? -0.5 : 0.5));
        double t = (double)k;
        hi = x - t * ln2HI;
        lo = t * ln2LO;
        xr = hi - lo;
    } else {
        k = 0; hi = x; lo = 0.0; xr = x;
    }
    double t = xr * xr;
    double c = xr - t * fma(t, fma(t, fma(t, fma(t, P5, P4), P3), P2), P1);
    double y = 1.0 - ((lo - (xr * c) / (2.0 - c)) - hi);
    if (k == 0) return y;
    if (k >= -1021) {
        if (k == 1024) return y * 2.0 * 0x1p1023;
        return u2d(d2u(y) + ((uint64_t)(int64_t)k << 52));
    }
    return u2d(d2u(y) + ((uint64_t)(int64_t)(k + 1000) << 52)) * 0x1p-1000;
}

/* lgamma(x), x > 0: shift to y >= 8 with the recurrence, then an 8-term Stirling series. */
double oracle_lgamma(double x)
{
    static const double HALF_LOG_2PI = 0.91893853320467274178,
        S0 = 8.3333333333333329e-02, S1 = -2.7777777777777779e-03, S2 = 7.9365079365079365e-04,
        S3 = -5.9523809523809529e-04, S4 = 8.4175084175084171e-04, S5 = -1.9175269175269176e-03,
        S6 = 6.4102564102564100e-03, S7 = -2.9550653594771242e-02;
    if (x != x) return x;
    if (x < 0.0) return NAN;
    if (x == 0.0) return INFINITY;
    if (x > 1.7e308) return INFINITY;
    double y = x, corr = 0.0;
    if (x < 8.0) {
        double prod = x;
        prod = prod * (x + 1.0); prod = prod * (x + 2.0); prod = prod * (x + 3.0);
        prod = prod * (x + 4.0); prod = prod * (x + 5.0); prod = prod * (x + 6.0);
        prod = prod * (x + 7.0);
        corr = oracle_log(prod);
        y = x + 8.0;
    }
    double r = 1.0 / y, r2 = r * r;
    double ser = r * fma(r2, fma(r2, fma(r2, fma(r2, fma(r2, fma(r2, fma(r2, S7, S6), S5), S4), S3), S2), S1), S0);
    double lg = ((y - 0.5) * oracle_log(y) - y) + HALF_LOG_2PI + ser;
    return lg - corr;
}

/* ------------------------------------------------------------------ samplers */
/* Stirling tail  log(k!) - [ (k+1/2) log(k+1) - (k+1) + log(2 pi)/2 ]  used by BTRS. */
static double btrs_fc(double k)
{
    static const double tab[10] = {
        0.08106146679532726, 0.04134069595540929, 0.02767792568499834, 0.02079067210376509,
        0.01664469118982119, 0.01387612882307075, 0.01189670994589177, 0.01041126526197209,
        0.009255462182712733, 0.008330563433362871 };
    if (k <= 9.0) return tab[(int)k];
    double ri = 1.0 / (k + 1.0), ri2 = ri * ri;
    return (1.0 / 12.0 - (1.0 / 360.0 - (1.0 / 1260.0) * ri2) * ri2) * ri;
}

/* Binomial(n, pp), n >= 1, 0 < pp < 1.  Small n*p: sequential-search inversion ("BINV",
 * Kachitvichyanukul & Schmeiser 1988) started from q^n.  Otherwise: Hormann's BTRS
 * ("The generation of binomial random variates", J. Stat. Comput. Simul. 46, 1993). */
static int binomial_draw(const ustream *s, int kstep, int n, double pp)
{
    int flip = pp > 0.5;
    double p = flip ? 1.0 - pp : pp;
    double q = 1.0 - p;
    double nd = (double)n;
    int x;
    if (nd * p < BINV_MAX) {
        double r = 1.0, b = q;
        unsigned e = (unsigned)n;
        for (;;) {
            if (e & 1u) r = r * b;
            e >>= 1;
            if (!e) break;
            b = b * b;
        }
        double uq[4];
        us_quad(s, (uint32_t)kstep >> 2, uq);
        double u = uq[kstep & 3];
        x = 0;
        if (u >= r) {
            double sr = p / q;
            double a = sr * (nd + 1.0);              /* sr*(n-x+1)/x = a/x - sr */
            double xd = 0.0;
            while (u >= r && x < n) {
                u = u - r;
                x += 1; xd += 1.0;
                r = r * fma(a, 1.0 / xd, -sr);
            }
        }
    } else {
        double spq = sqrt(nd * p * q);
        double b = 1.15 + 2.53 * spq;
        double a = -0.0873 + 0.0248 * b + 0.01 * p;
        double c = nd * p + 0.5;
        double vr = 0.92 - 4.2 / b;
        double alpha = (2.83 + 5.1 / b) * spq;
        double r = p / q;
        double m = floor((nd + 1.0) * p);
        double kd = 0.0;
        x = -1;
        for (uint32_t t = 0; t < 256; t++) {
            double uq[4];
            us_quad(s, BTRS_BLOCK0 + ((uint32_t)kstep << 8) + t, uq);
            double U = uq[0] - 0.5, V = uq[1];
            double us = 0.5 - fabs(U);
            kd = floor((2.0 * a / us + b) * U + c);
            if (us >= 0.07 && V <= vr) { x = (int)kd; break; }
            if (kd < 0.0 || kd > nd) continue;
            double lv = oracle_log(V * alpha / (a / (us * us) + b));
            double ub = (m + 0.5) * oracle_log((m + 1.0) / (r * (nd - m + 1.0)))
                      + (nd + 1.0) * oracle_log((nd - m + 1.0) / (nd - kd + 1.0))
                      + (kd + 0.5) * oracle_log(r * (nd - kd + 1.0) / (kd + 1.0))
                      + btrs_fc(m) + btrs_fc(nd - m) - btrs_fc(kd) - btrs_fc(nd - kd);
            if (lv <= ub) { x = (int)kd; break; }
        }
        if (x < 0) x = (kd < 0.0) ? 0 : (kd > nd) ? n : (int)kd;     /* 256 rejections: never in practice */
    }
    return flip ? n - x : x;
}

int oracle_binomial(uint64_t seed, uint32_t sweep, uint32_t stream, uint32_t elem, int kstep, int n, double pp)
{
    ustream s; us_init(&s, seed, sweep, stream, elem);
    return binomial_draw(&s, kstep, n, pp);
}

/* Standard normal by the ziggurat (Marsaglia & Tsang 2000; layer index independent of the value as in Doornik 2005),
 * 128 layers of equal area with right edges zig_x[0..128] (zig_x[0] the base strip's pseudo-edge, zig_x[1] = R).
 * Candidate from four Philox words: |u| = the 52 bits (w1 >> 12 : w0) as a fraction, sign = bit 7 of w1, layer = the
 * low 7 bits of w1; accepted at once when |u| < zig_x[i+1] / zig_x[i] (97 %).  Otherwise the base strip draws from
 * the tail beyond R (Marsaglia's exponential rejection) and the other layers test the wedge under the density.
 * Attempt t of a gamma draw takes its first candidate from words 0,1 of block t; every further pair of uniforms or
 * candidate takes the next block of the sequence NORM_BLOCK0 + (t << 8) + q, q = 0, 1, ... */
static const double zig_x[129] = SG_ZIG_X, zig_ratio[128] = SG_ZIG_RATIO;
#define NORM_BLOCK0 0x10000u

static double normal_draw(const ustream *s, uint32_t t, const uint32_t w_first[4])
{
    uint32_t w[4] = { w_first[0], w_first[1], w_first[2], w_first[3] };
    uint32_t q = 0;
    for (;;) {
        double mag = u2d(((uint64_t)(0x3ff00000u | (w[1] >> 12)) << 32) | w[0]) - 1.0;
        int i = (int)(w[1] & 127u), neg = (int)((w[1] >> 7) & 1u);
        double xi = zig_x[i];
        if (mag < zig_ratio[i]) { double x = mag * xi; return neg ? -x : x; }
        if (i == 0) {
            double xx, yy, u1, u2;
            do {
                us_block(s, NORM_BLOCK0 + (t << 8) + q, &u1, &u2); q++;
                xx = oracle_log(u1) / SG_ZIG_R;
                yy = oracle_log(u2);
            } while (-2.0 * yy < xx * xx);
            return neg ? xx - SG_ZIG_R : SG_ZIG_R - xx;
        }
        double x = mag * xi, xi1 = zig_x[i + 1];
        double f0 = oracle_exp(-0.5 * (xi * xi - x * x)), f1 = oracle_exp(-0.5 * (xi1 * xi1 - x * x));
        double ua, ub;
        us_block(s, NORM_BLOCK0 + (t << 8) + q, &ua, &ub); q++;
        if (f1 + ua * (f0 - f1) < 1.0) return neg ? -x : x;
        uint32_t ctr[4] = { NORM_BLOCK0 + (t << 8) + q, s->elem, s->stream, s->sweep };
        oracle_philox(ctr, s->key, w); q++;
    }
}

double oracle_normal(uint64_t seed, uint32_t sweep, uint32_t stream, uint32_t elem, uint32_t t)
{
    ustream s; us_init(&s, seed, sweep, stream, elem);
    uint32_t ctr[4] = { t, s.elem, s.stream, s.sweep }, w[4];
    oracle_philox(ctr, s.key, w);
    return normal_draw(&s, t, w);
}

/* Gamma(shape, rate) with the clamps of one_gamma_dist (src/gibbs_2.cpp:39-43);
 * Marsaglia & Tsang, "A simple method for generating gamma variables" (ACM TOMS 26, 2000).
 * Attempt t: normal from block t (words 0,1 first, see normal_draw), acceptance uniform u53(words 2,3) of block t. */
static double gamma_core(const ustream *s, double a, double scale)
{
    int boost = a < 1.0;
    double aa = boost ? a + 1.0 : a;
    double d = aa - 1.0 / 3.0;
    double c = 1.0 / sqrt(9.0 * d);
    double v = 1.0;
    for (uint32_t t = 0; t < 256; t++) {
        uint32_t ctr[4] = { t, s->elem, s->stream, s->sweep }, w[4];
        oracle_philox(ctr, s->key, w);
        double u2 = oracle_u53(w[2], w[3]);
        double x = normal_draw(s, t, w);
        v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        double x2 = x * x;
        if (u2 < 1.0 - 0.0331 * (x2 * x2)) break;
        if (oracle_log(u2) < 0.5 * x2 + d * (1.0 - v + oracle_log(v))) break;
    }
    double g = d * v;
    if (boost) {
        double ub, um;
        us_block(s, AUX_BLOCK, &ub, &um);
        g = g * oracle_exp(oracle_log(ub) / a);
    }
    return g * scale;
}

static double gamma_draw(const ustream *s, double shape, double rate)
{
    double a = (TINY < shape) ? shape : TINY;             /* one_gamma_dist, src/gibbs_2.cpp:40-41 */
    double invr = 1.0 / rate;
    double scale = (TINY < invr) ? invr : TINY;
    return gamma_core(s, a, scale);
}

/* the R::rgamma(shape, scale) the reference calls after its own clamps (src/gibbs_2.cpp:42) */
double oracle_rgamma(uint64_t seed, uint32_t sweep, uint32_t stream, uint32_t elem, double shape, double scale)
{
    ustream s; us_init(&s, seed, sweep, stream, elem);
    return gamma_core(&s, shape, scale);
}

/* the MH acceptance uniform of an element (second half of the auxiliary block) */
double oracle_mh_uniform(uint64_t seed, uint32_t sweep, uint32_t stream, uint32_t elem)
{
    ustream s; us_init(&s, seed, sweep, stream, elem);
    double ub, um;
    us_block(&s, AUX_BLOCK, &ub, &um);
    return um;
}

double oracle_gamma(uint64_t seed, uint32_t sweep, uint32_t stream, uint32_t elem, double shape, double rate)
{
    ustream s; us_init(&s, seed, sweep, stream, elem);
    return gamma_draw(&s, shape, rate);
}

/* Random scan order (src/gibbs_2.cpp:303-306: flist = arma::shuffle(flist) every sweep).
 * Fisher-Yates on {1..7} driven by six Philox words of stream 0. */
void oracle_perm(uint64_t seed, uint32_t sweep, int out[7])
{
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) }, w[8];
    uint32_t c0[4] = { 0, 0, ST_PERM, sweep }, c1[4] = { 1, 0, ST_PERM, sweep };
    oracle_philox(c0, key, w); oracle_philox(c1, key, w + 4);
    for (int i = 0; i < 7; i++) out[i] = i + 1;
    for (int idx = 6; idx >= 1; idx--) {
        uint32_t j = (uint32_t)(((uint64_t)w[6 - idx] * (uint64_t)(idx + 1)) >> 32);
        int t = out[idx]; out[idx] = out[j]; out[j] = t;
    }
}

/* ------------------------------------------------------------------ chain state (reference layouts)
 * All matrices column-major doubles as in Armadillo/R: M,W I x J; P,Ap,Bp I x K; E,Ae,Be K x J;
 * Z cube I x J x K with element (s,g,m) at s + I*(g + J*m). */
typedef struct {
    int I, J, K;
    const double *M, *W;
    double *Z;                        /* materialised cube, as the reference */
    double *P, *E, *Ap, *Bp, *Ae, *Be;
    double ap, bp, ae, be, lp, le, var_ap, var_ae;
    uint64_t seed;
    int nshards;
} chain;

#define ZIDX(c, s, g, m) ((size_t)(s) + (size_t)(c)->I * ((size_t)(g) + (size_t)(c)->J * (size_t)(m)))

/* One cell of step 1: Z[s,g,.] ~ Multinomial(trunc(M[s,g]), P[s,.] o E[.,g] / sum).  Counts above
 * NDIRECT are drawn as R's rmultinom does (sequential conditional binomials in k order;
 * zero-probability skip; pp<1 test; early exit at n<=0; remainder to the last class); smaller
 * counts by n categorical draws (same distribution, cost O(n log K) instead of O(K)).  z has stride zs.  src/gibbs_2.cpp:94-101. */
static void multinomial_cell(uint64_t seed, uint32_t sweep, uint32_t stream, int I, int K,
                             int s, int g, double Mval, const double *P, const double *E,
                             double *z, size_t zs)
{
    for (int m = 0; m < K; m++) z[m * zs] = 0.0;
    int n = (int)Mval;                                  /* int size (src/gibbs_2.cpp:31) */
    if (n <= 0) return;
    double S = 0.0;
    for (int m = 0; m < K; m++) S = fma(P[s + (size_t)I * m], E[m + (size_t)K * g], S);
    if (!(S > 0.0) || !(S <= 1.7976931348623157e308)) {  /* degenerate: everything to the last class */
        z[(K - 1) * zs] = (double)n;
        return;
    }
    ustream us; us_init(&us, seed, sweep, stream, (uint32_t)(s + I * g));
    if (n <= NDIRECT) {
        /* per-count algorithm switch: small counts use the direct method -- n categorical draws,
         * draw t from word (t&3) of block (t>>2): class = first k < K-1 with u*S < cumulative sum, else K-1 */
        for (int t = 0; t < n; t++) {
            double uq[4];
            us_quad(&us, (uint32_t)t >> 2, uq);
            double x = uq[t & 3] * S;
            double c = 0.0;
            int cls = K - 1;
            for (int m = 0; m < K - 1; m++) {
                c = c + P[s + (size_t)I * m] * E[m + (size_t)K * g];
                if (x < c) { cls = m; break; }
            }
            z[cls * zs] += 1.0;
        }
        return;
    }
    /* Large counts: sequential conditional binomials as in R's rmultinom (zero-probability skip,
     * pp<1 test, remainder to the class visited last), with two changes that leave the multinomial
     * law untouched and suit a GPU warp:
     *  (1) visiting order: classes sorted by the binary order of magnitude of P*E relative to S,
     *      bucket d = min(7, expo(S) - expo(P*E)) ascending, ties lowest index first (heavy classes
     *      first: the lanes of a warp meet their expensive draws together and the count is used up
     *      after the few classes that carry the mass);
     *  (2) the chain runs only while the remaining count exceeds NDIRECT; what is left is dealt to
     *      the classes not yet visited by the direct method (draw d: word d&3 of block
     *      TAIL_BLOCK0 + (d>>2); cumulative sums c_k in index order with the visited classes' P*E
     *      set to zero, Sr = c_{K-1}; x = u*Sr; class = first k < K-1 with x < c_k, else K-1).
     * Given the counts of the visited classes the rest is Multinomial(remaining, P*E of the rest). */
    double prob[ORACLE_KMAX];
    int order[ORACLE_KMAX];
    int eS = (int)((d2u(S) >> 52) & 0x7ff);
    int npos = 0;
    for (int m = 0; m < K; m++) prob[m] = P[s + (size_t)I * m] * E[m + (size_t)K * g];
    for (int d = 0; d < 8; d++)
        for (int m = 0; m < K; m++) {
            int dm = eS - (int)((d2u(prob[m]) >> 52) & 0x7ff);
            if (dm > 7) dm = 7;
            if (dm < 0) dm = 0;
            if (dm == d) order[npos++] = m;
        }
    double rem = S;
    int t = 0;
    while (n > NDIRECT && t < K - 1) {
        int cur = order[t];
        double best = prob[cur];
        if (best != 0.0) {
            double pp = best / rem;
            int zz = (pp > 0.0 && pp < 1.0) ? binomial_draw(&us, t, n, pp) : n;
            z[cur * zs] = (double)zz;
            n -= zz;
        }
        rem = rem - best;
        t++;
    }
    if (n <= 0) return;
    if (t == K - 1) { z[order[K - 1] * zs] += (double)n; return; }
    for (int j = 0; j < t; j++) prob[order[j]] = 0.0;           /* visited classes take no more */
    double cum[ORACLE_KMAX];
    double c = 0.0;
    for (int m = 0; m < K; m++) { c = c + prob[m]; cum[m] = c; }
    double Sr = c;
    for (int d = 0; d < n; d++) {
        double uq[4];
        us_quad(&us, TAIL_BLOCK0 + ((uint32_t)d >> 2), uq);
        double x = uq[d & 3] * Sr;
        int cls = K - 1;
        for (int m = 0; m < K - 1; m++) if (x < cum[m]) { cls = m; break; }
        z[cls * zs] += 1.0;
    }
}

/* step 1 -- src/gibbs_2.cpp:85-103 */
static void step1(chain *c, uint32_t sweep, uint32_t stream)
{
    for (int g = 0; g < c->J; g++)
        for (int s = 0; s < c->I; s++)
            multinomial_cell(c->seed, sweep, stream, c->I, c->K, s, g, c->M[s + (size_t)c->I * g],
                             c->P, c->E, c->Z + ZIDX(c, s, g, 0), (size_t)c->I * c->J);
}

/* cube sums -- src/gibbs_2.cpp:9-28 (slice-innermost loops over the cube). */
static void cube_sum_j(const chain *c, double *Zin /* I x K */)
{
    memset(Zin, 0, sizeof(double) * c->I * c->K);
    for (int i = 0; i < c->I; i++) for (int j = 0; j < c->J; j++) for (int k = 0; k < c->K; k++)
        Zin[i + (size_t)c->I * k] += c->Z[ZIDX(c, i, j, k)];
}
static void cube_sum_i_t(const chain *c, double *Znj /* K x J */)
{
    memset(Znj, 0, sizeof(double) * c->K * c->J);
    for (int i = 0; i < c->I; i++) for (int j = 0; j < c->J; j++) for (int k = 0; k < c->K; k++)
        Znj[k + (size_t)c->K * j] += c->Z[ZIDX(c, i, j, k)];
}

/* Ordered sum of a[0..n) in blocks of ORACLE_SUMBLK: every block of consecutive entries is added sequentially (from 0.0),
 * the block sums form the next level, until one value is left.  Overwrites a. */
static double blocked_sum(double *a, int n)
{
    if (n <= 0) return 0.0;
    for (;;) {
        int nb = (n + ORACLE_SUMBLK - 1) / ORACLE_SUMBLK;
        for (int b = 0; b < nb; b++) {
            int t1 = (b + 1) * ORACLE_SUMBLK; if (t1 > n) t1 = n;
            double sacc = 0.0;
            for (int t = b * ORACLE_SUMBLK; t < t1; t++) sacc = sacc + a[t];
            a[b] = sacc;
        }
        if (nb == 1) return a[0];
        n = nb;
    }
}

/* corrE = W * E.t()  (src/gibbs_2.cpp:111) with the specified summation order: fma chains over segments of
 * ORACLE_SEG genomes; per shard the segment sums are added by blocked_sum; the shard sums likewise. */
static void corr_E(const chain *c, double *corrE /* I x K */)
{
    int I = c->I, J = c->J, K = c->K;
    int nseg = (J + ORACLE_SEG - 1) / ORACLE_SEG;
    int nsh = c->nshards < 1 ? 1 : c->nshards;
    double *part = malloc(sizeof(double) * (nseg + 1)), *tot = malloc(sizeof(double) * (nsh + 1));
    for (int i = 0; i < I; i++) for (int k = 0; k < K; k++) {
        for (int sh = 0; sh < nsh; sh++) {
            /* shard sh owns segments [sh*nseg/nsh, (sh+1)*nseg/nsh) */
            int s0 = (int)((long long)sh * nseg / nsh), s1 = (int)((long long)(sh + 1) * nseg / nsh);
            int np = 0;
            for (int sg = s0; sg < s1; sg++) {
                int j0 = sg * ORACLE_SEG, j1 = j0 + ORACLE_SEG; if (j1 > J) j1 = J;
                double acc = 0.0;
                for (int j = j0; j < j1; j++) acc = fma(c->W[i + (size_t)I * j], c->E[k + (size_t)K * j], acc);
                part[np++] = acc;
            }
            tot[sh] = blocked_sum(part, np);
        }
        corrE[i + (size_t)I * k] = blocked_sum(tot, nsh);
    }
    free(part); free(tot);
}

/* step 2 -- src/gibbs_2.cpp:106-120 */
static void step2(chain *c, uint32_t sweep)
{
    int I = c->I, K = c->K;
    double *corrE = malloc(sizeof(double) * I * K), *Zin = malloc(sizeof(double) * I * K);
    corr_E(c, corrE); cube_sum_j(c, Zin);
    for (int s = 0; s < I; s++) for (int m = 0; m < K; m++) {
        size_t e = s + (size_t)I * m;
        ustream us; us_init(&us, c->seed, sweep, ST_P, (uint32_t)e);
        double shape = c->Ap[e] + 1 + Zin[e];
        double rate = c->Bp[e] + corrE[e];
        c->P[e] = gamma_draw(&us, shape, rate);
    }
    free(corrE); free(Zin);
}

/* step 3 -- src/gibbs_2.cpp:123-137 ; corrP = P.t() * W as an fma chain over i. */
static void step3(chain *c, uint32_t sweep)
{
    int I = c->I, J = c->J, K = c->K;
    double *Znj = malloc(sizeof(double) * K * J);
    cube_sum_i_t(c, Znj);
    for (int g = 0; g < J; g++) for (int m = 0; m < K; m++) {
        size_t e = m + (size_t)K * g;
        double corrP = 0.0;
        for (int i = 0; i < I; i++) corrP = fma(c->P[i + (size_t)I * m], c->W[i + (size_t)I * g], corrP);
        ustream us; us_init(&us, c->seed, sweep, ST_E, (uint32_t)e);
        double shape = c->Ae[e] + 1 + Znj[e];
        double rate = c->Be[e] + corrP;
        c->E[e] = gamma_draw(&us, shape, rate);
    }
    free(Znj);
}

/* steps 4,5 -- src/gibbs_2.cpp:140-169 */
static void step_b(double *B, const double *X, const double *A, size_t n, double a, double b,
                   uint64_t seed, uint32_t sweep, uint32_t stream)
{
    for (size_t e = 0; e < n; e++) {
        ustream us; us_init(&us, seed, sweep, stream, (uint32_t)e);
        double shape = A[e] + a + 1;
        double rate = X[e] + b;
        double g = gamma_draw(&us, shape, rate);
        B[e] = (TINY < g) ? g : TINY;
    }
}

/* steps 6,7 -- src/gibbs_2.cpp:172-251 */
static void step_a(double *A, const double *X, const double *B, size_t n, double l, double var,
                   uint64_t seed, uint32_t sweep, uint32_t stream)
{
    for (size_t e = 0; e < n; e++) {
        ustream us; us_init(&us, seed, sweep, stream, (uint32_t)e);
        double a = A[e];
        double shape = (a * a) / var;
        double rate = a / var;
        double g = gamma_draw(&us, shape, rate);
        double y = (TINY < g) ? g : TINY;
        double lnp = (y - a) * (oracle_log(B[e]) + oracle_log(X[e]) - l)
                   + oracle_lgamma(a + 1) - oracle_lgamma(y + 1)
                   + oracle_lgamma((a * a) / var) - oracle_lgamma((y * y) / var)
                   + (((y * y) - (a * a)) / var) * oracle_log((a * y) / var)
                   + oracle_log(y) - oracle_log(a);
        double pch = oracle_exp(lnp);
        if (X[e] == 0) pch = (y < a) ? 1.0 : 0.0;
        double ub, um;
        us_block(&us, AUX_BLOCK, &ub, &um);
        if (um <= pch) A[e] = y;
    }
}

/* per-sample log-likelihood -- R/cpp_eBayesNMF.R:180-185, specified summation order. */
double oracle_loglik(int I, int J, int K, const double *M, const double *W, const double *P, const double *E)
{
    int n2 = 1; while (n2 < J) n2 <<= 1;
    double *col = calloc(n2, sizeof(double)), *lgm = calloc(n2, sizeof(double));
    for (int j = 0; j < J; j++) {
        double a = 0.0, b = 0.0;
        for (int c0 = 0; c0 < I; c0 += ORACLE_LLCHUNK) {            /* 32-row chunks, added in order */
            int c1 = c0 + ORACLE_LLCHUNK; if (c1 > I) c1 = I;
            double ac = 0.0, bc = 0.0;
            for (int i = c0; i < c1; i++) {
                double pe = 0.0;
                for (int k = 0; k < K; k++) pe = fma(P[i + (size_t)I * k], E[k + (size_t)K * j], pe);
                double lam = pe * W[i + (size_t)I * j];
                double m = (double)(int)M[i + (size_t)I * j];
                ac = ac + (m * oracle_log(lam) - lam);
                bc = bc + oracle_lgamma(m + 1.0);
            }
            a = a + ac; b = b + bc;
        }
        col[j] = a; lgm[j] = b;
    }
    for (int s = n2 / 2; s >= 1; s >>= 1)
        for (int t = 0; t < s; t++) { col[t] = col[t] + col[t + s]; lgm[t] = lgm[t] + lgm[t + s]; }
    double ll = col[0] - lgm[0];
    free(col); free(lgm);
    return ll;
}

/* ------------------------------------------------------------------ driver: src/gibbs_2.cpp:253-424
 * Only the flag set reachable from the package (Zfixed=Thetafixed=Afixed=FALSE); Pfixed and
 * keep_par vary.  State arrays are updated in place (the R wrapper passes copies).
 * If Zin is NULL the initial Z is drawn with stream 8 (sweep = sweep0), replacing the R loop
 * R/cpp_eBayesNMF.R:79-89.  Outputs may be NULL.  forced_order (7 ints per sweep, 0 = skip)
 * overrides the random scan for teacher-forced tests. */
int oracle_gibbs_run(int I, int J, int K, const double *M, const double *W, const double *Zinit,
                     double *P, double *E, double *Ap, double *Bp, double *Ae, double *Be,
                     const double hyper[8], int burn, int eval, int Pfixed, int keep_par,
                     uint64_t seed, uint32_t sweep0, int nshards, const int *forced_order,
                     double *Ps, double *Es, double *Aps, double *Bps, double *Aes, double *Bes,
                     double *Zlast, double *loglik, double *Zin_out, double *Znj_out)
{
    chain c = { I, J, K, M, W, NULL, P, E, Ap, Bp, Ae, Be,
                hyper[0], hyper[1], hyper[2], hyper[3], hyper[4], hyper[5], hyper[6], hyper[7], seed, nshards };
    size_t nz = (size_t)I * J * K, nP = (size_t)I * K, nE = (size_t)K * J;
    c.Z = malloc(sizeof(double) * nz);
    if (!c.Z) return 4;
    if (Zinit) memcpy(c.Z, Zinit, sizeof(double) * nz);
    else step1(&c, sweep0, ST_ZINIT);
    for (int k = 0; k < burn + eval; k++) {
        uint32_t sweep = sweep0 + (uint32_t)k;
        int order[7];
        if (forced_order) memcpy(order, forced_order + 7 * (size_t)k, sizeof(order));
        else oracle_perm(seed, sweep, order);
        for (int f = 0; f < 7; f++) {
            switch (order[f]) {
            case 1: step1(&c, sweep, ST_Z); break;
            case 2: if (!Pfixed) step2(&c, sweep); break;
            case 3: step3(&c, sweep); break;
            case 4: if (!Pfixed) step_b(Bp, P, Ap, nP, c.ap, c.bp, seed, sweep, ST_BP); break;
            case 5: step_b(Be, E, Ae, nE, c.ae, c.be, seed, sweep, ST_BE); break;
            case 6: if (!Pfixed) step_a(Ap, P, Bp, nP, c.lp, c.var_ap, seed, sweep, ST_AP); break;
            case 7: step_a(Ae, E, Be, nE, c.le, c.var_ae, seed, sweep, ST_AE); break;
            default: break;
            }
        }
        if (k >= burn) {
            int r = k - burn;
            if (Ps) memcpy(Ps + nP * r, P, sizeof(double) * nP);
            if (Es) memcpy(Es + nE * r, E, sizeof(double) * nE);
            if (keep_par && Aps) memcpy(Aps + nP * r, Ap, sizeof(double) * nP);
            if (keep_par && Aes) memcpy(Aes + nE * r, Ae, sizeof(double) * nE);
            if (Bps) memcpy(Bps + nP * r, Bp, sizeof(double) * nP);
            if (Bes) memcpy(Bes + nE * r, Be, sizeof(double) * nE);
            if (loglik) loglik[r] = oracle_loglik(I, J, K, M, W, P, E);
        }
    }
    if (Zlast) memcpy(Zlast, c.Z, sizeof(double) * nz);
    if (Zin_out) cube_sum_j(&c, Zin_out);
    if (Znj_out) cube_sum_i_t(&c, Znj_out);
    free(c.Z);
    return 0;
}

/* R's rmultinom LITERALLY (nmath/rmultinom.c as SURVEY.md Appendix C restates it): classes in index order, no re-ordering,
 * no early stop, no direct method -- for k < K-1: skip zero probabilities; pp = prob[k] / p_tot; rN[k] = pp < 1 ?
 * rbinom(n, pp) : n; n -= rN[k]; stop when n <= 0; p_tot -= prob[k]; the last class takes the remainder.  prob = the
 * reference's Fi = P[s,k] E[k,g] / (PE)[s,g] (src/gibbs_2.cpp:89-93).  The binomials are this file's binomial_draw on
 * the cell's stream (conditional binomial k takes word k&3 of block k>>2).  NOT the algorithm of the draw specification:
 * it exists so that the specification's re-ordered chain + direct remainder (and the CUDA kernel) can be cross-checked
 * statistically against the structure R actually uses. */
void oracle_multinomial_cell_literal(uint64_t seed, uint32_t sweep, uint32_t stream, int I, int K,
                                     int s, int g, double Mval, const double *P, const double *E, double *z /* K */)
{
    for (int m = 0; m < K; m++) z[m] = 0.0;
    int n = (int)Mval;
    if (n <= 0) return;
    double S = 0.0;
    for (int m = 0; m < K; m++) S = fma(P[s + (size_t)I * m], E[m + (size_t)K * g], S);
    if (!(S > 0.0) || !(S <= 1.7976931348623157e308)) { z[K - 1] = (double)n; return; }
    ustream us; us_init(&us, seed, sweep, stream, (uint32_t)(s + I * g));
    long double p_tot = 1.0L;
    for (int m = 0; m < K - 1; m++) {
        double prob = (P[s + (size_t)I * m] * E[m + (size_t)K * g]) / S;
        if (prob != 0.0) {
            double pp = (double)((long double)prob / p_tot);
            int zz = (pp < 1.0) ? ((pp > 0.0) ? binomial_draw(&us, m, n, pp) : 0) : n;
            z[m] = (double)zz;
            n -= zz;
        }
        if (n <= 0) return;
        p_tot -= (long double)prob;
    }
    z[K - 1] = (double)n;
}

/* teacher-forced single cell for unit tests */
void oracle_multinomial_cell(uint64_t seed, uint32_t sweep, uint32_t stream, int I, int K,
                             int s, int g, double Mval, const double *P, const double *E, double *z /* K */)
{
    multinomial_cell(seed, sweep, stream, I, K, s, g, Mval, P, E, z, 1);
}
