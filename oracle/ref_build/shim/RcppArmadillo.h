// Stand-in for <RcppArmadillo.h> -- TEST INFRASTRUCTURE ONLY (oracle/ref_build).
//
// The reference's sampler src/gibbs_2.cpp is compiled UNMODIFIED, from where it lies under
// /root/reference, against this header instead of Rcpp / RcppArmadillo / R (none of which exist in
// this image).  The header supplies exactly the slice of those libraries the file uses:
//   * arma::mat / cube / rowvec / Col<int> / Row<int> with eager element-wise arithmetic
//     (each operator is one correctly rounded fp64 operation per element, evaluated in C++'s
//     left-to-right order like Armadillo's expression templates evaluate per element);
//   * matrix products with the summation orders of the draw specification (W*E.t(): 32-column
//     segment fma chains, blocked sums; P.t()*W and P*E: fma chains);
//   * R::rmultinom, R::rgamma, arma::shuffle, arma::randu, arma::log/exp/lgamma routed to the
//     draw specification (oracle/gibbs_oracle.c) through a replay context that addresses every
//     call by (sweep, step, element) exactly as the reference's loops visit them;
//   * Rcpp::NumericVector / IntegerVector / List / wrap / named arguments, enough to build the
//     returned list.
// What this validates: the reference's own update equations, clamps, loop orders, flag logic,
// storage rules and return list against the oracle's restatement, bit for bit.  What it cannot
// validate: R's nmath and Armadillo's RNG themselves (absent) -- they are replaced here.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <initializer_list>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

extern "C" {
double oracle_log(double), oracle_exp(double), oracle_lgamma(double);
double oracle_rgamma(uint64_t seed, uint32_t sweep, uint32_t stream, uint32_t elem, double shape, double scale);
double oracle_mh_uniform(uint64_t seed, uint32_t sweep, uint32_t stream, uint32_t elem);
void oracle_perm(uint64_t seed, uint32_t sweep, int out[7]);
void oracle_multinomial_cell(uint64_t seed, uint32_t sweep, uint32_t stream, int I, int K, int s, int g, double Mval,
                             const double *P, const double *E, double *z);
}

// ------------------------------------------------------------------ replay context
namespace refctx {
struct Ctx {
    uint64_t seed = 0; uint32_t sweep0 = 0; int I = 0, J = 0, K = 0; bool Pfixed = false;
    std::vector<double> P, E;            // mirror of the chain's P and E (they are direct rgamma results)
    int sweep = -1;                      // current sweep (incremented by shuffle)
    int order[7] = {0, 0, 0, 0, 0, 0, 0};
    int pos = 0;                         // index into order of the step being executed
    long calls = 0;                      // primitive calls made inside the current step
    double max_prob_dev = 0.0;           // max |Fi - P*E/S| seen by rmultinom (input equivalence check)
    int active(int s) const { return !((s == 2 || s == 4 || s == 6) && Pfixed); }
    long ncalls(int s) const { return s == 1 ? (long)I * J : (s == 2 || s == 4 || s == 6) ? (long)I * K : (long)K * J; }
    // the step the next primitive call belongs to (skipping steps the guards of :310-316 skip)
    int step() {
        while (pos < 7 && (!active(order[pos]) || calls >= ncalls(order[pos]))) { if (active(order[pos]) && calls >= ncalls(order[pos])) calls = 0; ++pos; }
        if (pos >= 7) throw std::runtime_error("refctx: primitive call outside a sweep");
        return order[pos];
    }
};
inline Ctx &ctx() { static Ctx c; return c; }
} // namespace refctx

// ------------------------------------------------------------------ arma
namespace arma {
namespace fill { struct zeros_t {}; static const zeros_t zeros{}; }
typedef std::size_t uword;

struct mat;
struct trans_view { const mat &m; };

struct mat {
    uword n_rows = 0, n_cols = 0;
    std::vector<double> d;
    mat() {}
    mat(uword r, uword c) : n_rows(r), n_cols(c), d(r * c, 0.0) {}
    mat(uword r, uword c, fill::zeros_t) : n_rows(r), n_cols(c), d(r * c, 0.0) {}
    mat(const trans_view &t);                                    // arma::mat X = Y.t();
    double &operator()(uword i, uword j) { return d[i + n_rows * j]; }
    double operator()(uword i, uword j) const { return d[i + n_rows * j]; }
    uword size() const { return d.size(); }
    double *begin() { return d.data(); }
    double &operator[](uword i) { return d[i]; }
    double operator[](uword i) const { return d[i]; }
    mat col(uword c) const { mat r(n_rows, 1); for (uword i = 0; i < n_rows; i++) r.d[i] = (*this)(i, c); return r; }
    mat row(uword rr) const { mat r(1, n_cols); for (uword j = 0; j < n_cols; j++) r.d[j] = (*this)(rr, j); return r; }
    trans_view t() const { return trans_view{*this}; }
    void clear() { d.clear(); n_rows = n_cols = 0; }
};
inline mat::mat(const trans_view &t) : n_rows(t.m.n_cols), n_cols(t.m.n_rows), d(t.m.d.size()) {
    for (uword i = 0; i < t.m.n_rows; i++) for (uword j = 0; j < t.m.n_cols; j++) d[j + n_rows * i] = t.m(i, j);
}
struct rowvec : mat {
    rowvec() {}
    rowvec(uword n, fill::zeros_t) : mat(1, n) {}
    rowvec(const mat &m) : mat(m) {}
};

inline void same(const mat &a, const mat &b) { if (a.n_rows != b.n_rows || a.n_cols != b.n_cols) throw std::runtime_error("arma shim: size mismatch"); }
#define ARMA_BIN(op, expr) \
    inline mat operator op(const mat &a, const mat &b) { same(a, b); mat r(a.n_rows, a.n_cols); for (uword i = 0; i < r.d.size(); i++) { double x = a.d[i], y = b.d[i]; r.d[i] = expr; } return r; } \
    inline mat operator op(const mat &a, double y) { mat r(a.n_rows, a.n_cols); for (uword i = 0; i < r.d.size(); i++) { double x = a.d[i]; r.d[i] = expr; } return r; } \
    inline mat operator op(double x, const mat &b) { mat r(b.n_rows, b.n_cols); for (uword i = 0; i < r.d.size(); i++) { double y = b.d[i]; r.d[i] = expr; } return r; }
ARMA_BIN(+, x + y)
ARMA_BIN(-, x - y)
ARMA_BIN(/, x / y)
ARMA_BIN(%, x * y)
#undef ARMA_BIN
// matrix products.  Reduction index and order follow the draw specification.
inline mat operator*(const mat &a, const mat &b) {               // P*E, P.col(m)*E.row(m): fma chain over the inner index
    if (a.n_cols != b.n_rows) throw std::runtime_error("arma shim: product size mismatch");
    mat r(a.n_rows, b.n_cols);
    for (uword j = 0; j < b.n_cols; j++) for (uword i = 0; i < a.n_rows; i++) {
        double acc = 0.0;
        for (uword k = 0; k < a.n_cols; k++) acc = std::fma(a(i, k), b(k, j), acc);
        r(i, j) = acc;
    }
    return r;
}
inline mat operator*(const mat &a, const trans_view &bt) {       // W * E.t(): fma chains over 32-column segments, segment sums added in blocks of 32
    const mat &b = bt.m;
    if (a.n_cols != b.n_cols) throw std::runtime_error("arma shim: product size mismatch");
    mat r(a.n_rows, b.n_rows);
    std::vector<double> part;
    for (uword i = 0; i < a.n_rows; i++) for (uword k = 0; k < b.n_rows; k++) {
        part.clear();
        for (uword j0 = 0; j0 < a.n_cols; j0 += 32) {
            uword j1 = j0 + 32 < a.n_cols ? j0 + 32 : a.n_cols;
            double acc = 0.0;
            for (uword j = j0; j < j1; j++) acc = std::fma(a(i, j), b(k, j), acc);
            part.push_back(acc);
        }
        uword n = part.size();
        while (n > 1) {                                          // draw specification v7: blocked ordered sum
            uword nb = (n + 31) / 32;
            for (uword bb = 0; bb < nb; bb++) {
                uword t1 = (bb + 1) * 32 < n ? (bb + 1) * 32 : n;
                double sacc = 0.0;
                for (uword t = bb * 32; t < t1; t++) sacc = sacc + part[t];
                part[bb] = sacc;
            }
            n = nb;
        }
        r(i, k) = 0.0 + (n ? 0.0 + part[0] : 0.0);
    }
    return r;
}
inline mat operator*(const trans_view &at, const mat &b) {       // P.t() * W: fma chain over i
    const mat &a = at.m;
    if (a.n_rows != b.n_rows) throw std::runtime_error("arma shim: product size mismatch");
    mat r(a.n_cols, b.n_cols);
    for (uword j = 0; j < b.n_cols; j++) for (uword k = 0; k < a.n_cols; k++) {
        double acc = 0.0;
        for (uword i = 0; i < a.n_rows; i++) acc = std::fma(a(i, k), b(i, j), acc);
        r(k, j) = acc;
    }
    return r;
}
inline mat trans(const mat &a) { mat r(a.n_cols, a.n_rows); for (uword i = 0; i < a.n_rows; i++) for (uword j = 0; j < a.n_cols; j++) r(j, i) = a(i, j); return r; }
#define ARMA_FUN(name, expr) inline mat name(const mat &a) { mat r(a.n_rows, a.n_cols); for (uword i = 0; i < r.d.size(); i++) { double x = a.d[i]; r.d[i] = expr; } return r; }
ARMA_FUN(log, oracle_log(x))
ARMA_FUN(exp, oracle_exp(x))
ARMA_FUN(lgamma, oracle_lgamma(x))
#undef ARMA_FUN
inline mat pow(const mat &a, int e) { if (e != 2) throw std::runtime_error("arma shim: pow exponent"); mat r(a.n_rows, a.n_cols); for (uword i = 0; i < r.d.size(); i++) r.d[i] = a.d[i] * a.d[i]; return r; }
inline double accu(const mat &a) { double s = 0; for (double v : a.d) s += v; return s; }
// U = arma::randu(r, c) in steps 6/7 (src/gibbs_2.cpp:193,234): the MH uniform of each element
inline mat randu(uword r, uword c) {
    refctx::Ctx &x = refctx::ctx();
    const int step = x.order[x.pos];                 // still the MH step whose proposals were just drawn
    mat u(r, c);
    for (uword e = 0; e < u.d.size(); e++) u.d[e] = oracle_mh_uniform(x.seed, x.sweep0 + (uint32_t)x.sweep, (uint32_t)step, (uint32_t)e);
    return u;
}

struct cube;
struct slice_ref { cube &c; uword s; slice_ref &operator=(const mat &m); };
struct cube {
    uword n_rows = 0, n_cols = 0, n_slices = 0;
    std::vector<double> d;
    cube() {}
    cube(uword r, uword c, uword s) : n_rows(r), n_cols(c), n_slices(s), d(r * c * s, 0.0) {}
    cube(uword r, uword c, uword s, fill::zeros_t) : n_rows(r), n_cols(c), n_slices(s), d(r * c * s, 0.0) {}
    double &operator()(uword i, uword j, uword k) { return d[i + n_rows * (j + n_cols * k)]; }
    double operator()(uword i, uword j, uword k) const { return d[i + n_rows * (j + n_cols * k)]; }
    slice_ref slice(uword s) { return slice_ref{*this, s}; }
    void clear() { d.clear(); n_rows = n_cols = n_slices = 0; }
};
inline slice_ref &slice_ref::operator=(const mat &m) {
    if (m.n_rows != c.n_rows || m.n_cols != c.n_cols) throw std::runtime_error("arma shim: slice size mismatch");
    for (uword i = 0; i < m.d.size(); i++) c.d[i + m.d.size() * s] = m.d[i];
    return *this;
}
#define CUBE_BIN(op, expr) \
    inline cube operator op(const cube &a, const cube &b) { cube r(a.n_rows, a.n_cols, a.n_slices); for (uword i = 0; i < r.d.size(); i++) { double x = a.d[i], y = b.d[i]; r.d[i] = expr; } return r; } \
    inline cube operator op(const cube &a, double y) { cube r(a.n_rows, a.n_cols, a.n_slices); for (uword i = 0; i < r.d.size(); i++) { double x = a.d[i]; r.d[i] = expr; } return r; }
CUBE_BIN(+, x + y)
CUBE_BIN(-, x - y)
CUBE_BIN(%, x * y)
#undef CUBE_BIN
inline cube log(const cube &a) { cube r(a.n_rows, a.n_cols, a.n_slices); for (uword i = 0; i < r.d.size(); i++) r.d[i] = oracle_log(a.d[i]); return r; }
inline cube lgamma(const cube &a) { cube r(a.n_rows, a.n_cols, a.n_slices); for (uword i = 0; i < r.d.size(); i++) r.d[i] = oracle_lgamma(a.d[i]); return r; }
inline double accu(const cube &a) { double s = 0; for (double v : a.d) s += v; return s; }

template <typename T> struct Col {
    std::vector<T> d;
    Col() {}
    explicit Col(uword n) : d(n) {}
    T *begin() { return d.data(); }
    T &operator[](uword i) { return d[i]; }
    T operator[](uword i) const { return d[i]; }
    uword size() const { return d.size(); }
};
template <typename T> struct Row {
    std::vector<T> d;
    Row() {}
    Row(std::initializer_list<T> l) : d(l) {}
    T &operator[](uword i) { return d[i]; }
    T operator[](uword i) const { return d[i]; }
    uword size() const { return d.size(); }
};
// flist = arma::shuffle(flist) at the top of every sweep (src/gibbs_2.cpp:306): the scan order of the draw specification
inline Row<int> shuffle(const Row<int> &) {
    refctx::Ctx &x = refctx::ctx();
    x.sweep += 1; x.pos = 0; x.calls = 0;
    oracle_perm(x.seed, x.sweep0 + (uint32_t)x.sweep, x.order);
    Row<int> r;
    r.d.assign(x.order, x.order + 7);
    return r;
}
} // namespace arma

// ------------------------------------------------------------------ R nmath entry points the file calls
namespace R {
// R::rmultinom(size, prob, K, rN) from one_multinom (src/gibbs_2.cpp:34); cells come in the order of :94-95 (s outer, g inner)
inline void rmultinom(int n, double *prob, int K, int *rN) {
    refctx::Ctx &x = refctx::ctx();
    const int step = x.step();
    if (step != 1) throw std::runtime_error("refctx: rmultinom outside step 1");
    const long c = x.calls++;
    const int s = (int)(c / x.J), g = (int)(c % x.J);
    // the probabilities the reference passes are P o E / (P E): check they are the ones the specification normalises itself
    double S = 0.0;
    for (int m = 0; m < K; m++) S = std::fma(x.P[s + (size_t)x.I * m], x.E[m + (size_t)K * g], S);
    for (int m = 0; m < K; m++) {
        const double q = x.P[s + (size_t)x.I * m] * x.E[m + (size_t)K * g] / S;
        const double dev = std::fabs(q - prob[m]);
        if (dev > x.max_prob_dev) x.max_prob_dev = dev;
    }
    std::vector<double> z(K);
    oracle_multinomial_cell(x.seed, x.sweep0 + (uint32_t)x.sweep, 1, x.I, K, s, g, (double)n, x.P.data(), x.E.data(), z.data());
    for (int m = 0; m < K; m++) rN[m] = (int)z[m];
}
// R::rgamma(shape, scale) from one_gamma_dist (src/gibbs_2.cpp:42); elements in the order of the step's loops
inline double rgamma(double shape, double scale) {
    refctx::Ctx &x = refctx::ctx();
    const int step = x.step();
    const long c = x.calls++;
    uint32_t e;
    if (step == 2 || step == 4 || step == 6) { const long s = c / x.K, m = c % x.K; e = (uint32_t)(s + (long)x.I * m); }      // for s: for m
    else if (step == 3 || step == 5 || step == 7) { const long m = c / x.J, g = c % x.J; e = (uint32_t)(m + (long)x.K * g); } // for m: for g
    else throw std::runtime_error("refctx: rgamma inside step 1");
    const double v = oracle_rgamma(x.seed, x.sweep0 + (uint32_t)x.sweep, (uint32_t)step, e, shape, scale);
    if (step == 2) x.P[e] = v;            // P(s,m) = one_gamma_dist(...)  (:117)
    if (step == 3) x.E[e] = v;            // E(m,g) = one_gamma_dist(...)  (:134)
    return v;
}
} // namespace R

// ------------------------------------------------------------------ Rcpp
#define NA_LOGICAL (-2147483647 - 1)
namespace Rcpp {
struct Named { std::string name; int value = 0; Named &operator=(int v) { value = v; return *this; } Named &operator=(std::size_t v) { value = (int)v; return *this; } };
struct NamedPlaceholder { Named operator[](const std::string &n) const { Named x; x.name = n; return x; } };
static const NamedPlaceholder _{};
struct IntegerVector {
    std::vector<int> d; std::vector<std::string> names;
    static IntegerVector create(const Named &a, const Named &b, const Named &c) { IntegerVector v; for (const Named *n : {&a, &b, &c}) { v.d.push_back(n->value); v.names.push_back(n->name); } return v; }
    static IntegerVector create(const Named &a, const Named &b, const Named &c, const Named &e) { IntegerVector v; for (const Named *n : {&a, &b, &c, &e}) { v.d.push_back(n->value); v.names.push_back(n->name); } return v; }
    int size() const { return (int)d.size(); }
    int operator[](int i) const { return d[i]; }
};
struct NumericVector;
struct AttrProxy { NumericVector &v; void operator=(const IntegerVector &dims); };
struct NumericVector {
    std::vector<double> d; IntegerVector dim; bool is_na = false;
    NumericVector() {}
    explicit NumericVector(std::size_t n) : d(n, 0.0) {}
    NumericVector(int na) : is_na(true) { (void)na; }             // NA_LOGICAL slot of the returned list
    double &operator[](std::size_t i) { return d[i]; }
    AttrProxy attr(const char *) { return AttrProxy{*this}; }
};
inline void AttrProxy::operator=(const IntegerVector &dims) { v.dim = dims; }
inline NumericVector wrap(const arma::cube &c) { NumericVector v(c.d.size()); v.d = c.d; return v; }
struct List {
    std::vector<NumericVector> items;
    template <typename... A> static List create(const A &...a) { List l; (l.items.push_back(NumericVector(a)), ...); return l; }
};
} // namespace Rcpp
