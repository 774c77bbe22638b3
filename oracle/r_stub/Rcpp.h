// Stand-in for <Rcpp.h> -- TEST INFRASTRUCTURE ONLY (oracle/r_stub).
//
// R, Rcpp and RcppArmadillo do not exist in this image, so the reference-side binding r_shim/signer_gibbs_shim.cpp
// (the file a signeR maintainer drops into src/ in place of gibbs_2.cpp) could never be
// compiled.  This header models exactly the slice of the R API / Rcpp the shim touches -- SEXP vectors with a "dim"
// attribute that carries names, NumericVector / NumericMatrix / IntegerVector / List, as<>(), RNGScope + unif_rand(),
// stop(), BEGIN_RCPP / END_RCPP -- so that the shim is compiled UNMODIFIED and its 24-SEXP entry point is executed on
// the GPU box by tests/test_gpu_rshim.py (through oracle/r_stub/driver.cpp), and its 8-slot list compared with the one
// the reference's own src/gibbs_2.cpp returns (oracle/_ref).  Nothing in signer_b200/ includes this file.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

typedef std::ptrdiff_t R_xlen_t;

struct SEXPREC {
    enum Type { NIL, REAL, INT, LGL, LIST, ERROR } type = NIL;
    std::vector<double> real;
    std::vector<int> ints;                 // INT / LGL payload
    std::vector<SEXPREC *> items;          // LIST payload
    SEXPREC *dim = nullptr;                // "dim" attribute (an INT vector, possibly with names)
    std::vector<std::string> names;        // names of an INT vector
    std::string msg;                       // ERROR payload (what R would raise)
};
typedef SEXPREC *SEXP;

namespace rstub {
inline std::deque<std::unique_ptr<SEXPREC>> &arena() { static std::deque<std::unique_ptr<SEXPREC>> a; return a; }
inline SEXP alloc(SEXPREC::Type t) { arena().emplace_back(new SEXPREC()); arena().back()->type = t; return arena().back().get(); }
inline void release_all() { arena().clear(); }
// R's RNG as far as the shim sees it: a 64-bit generator behind unif_rand(), seeded by the driver (the role of set.seed())
inline uint64_t &rng_state() { static uint64_t s = 1; return s; }
inline int &rng_scopes() { static int n = 0; return n; }
inline double unif() {                      // splitmix64 -> 32 random bits -> (0,1), like unif_rand's 32-bit resolution
    uint64_t z = (rng_state() += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; z ^= z >> 31;
    return ((double)(uint32_t)(z >> 32) + 0.5) / 4294967296.0;
}
} // namespace rstub

#define NA_LOGICAL (-2147483647 - 1)
inline double unif_rand() { if (rstub::rng_scopes() <= 0) throw std::runtime_error("unif_rand() outside an RNGScope"); return rstub::unif(); }

#define RcppExport extern "C"
#define BEGIN_RCPP try {
#define END_RCPP } catch (std::exception &ex_) { SEXP e_ = rstub::alloc(SEXPREC::ERROR); e_->msg = ex_.what(); return e_; }

namespace Rcpp {

struct RNGScope { RNGScope() { ++rstub::rng_scopes(); } ~RNGScope() { --rstub::rng_scopes(); } };

struct NamedInt { std::string name; int value; };
struct NamedProxy { std::string name; NamedInt operator=(int v) const { return NamedInt{name, v}; } };
struct NamedPlaceholder { NamedProxy operator[](const char *n) const { return NamedProxy{n}; } };
static const NamedPlaceholder _{};

class IntegerVector {
public:
    SEXP s;
    IntegerVector() : s(rstub::alloc(SEXPREC::INT)) {}
    template <typename... A> static IntegerVector create(const A &...a) { IntegerVector v; (v.push(a), ...); return v; }
    void push(const NamedInt &n) { s->ints.push_back(n.value); s->names.push_back(n.name); }
};

class NumericVector {
public:
    SEXP s;
    NumericVector() : s(rstub::alloc(SEXPREC::REAL)) {}
    explicit NumericVector(R_xlen_t n) : s(rstub::alloc(SEXPREC::REAL)) { s->real.assign((size_t)n, 0.0); }
    NumericVector(SEXP x) : s(x) { if (!x || x->type != SEXPREC::REAL) throw std::runtime_error("not a numeric vector"); }
    double *begin() { return s->real.data(); }
    R_xlen_t size() const { return (R_xlen_t)s->real.size(); }
    struct AttrProxy { SEXP s; void operator=(const IntegerVector &d) { s->dim = d.s; } };
    AttrProxy attr(const char *name) { if (std::strcmp(name, "dim")) throw std::runtime_error("stub: only the dim attribute"); return AttrProxy{s}; }
    operator SEXP() const { return s; }
};

class NumericMatrix {
public:
    SEXP s;
    NumericMatrix(SEXP x) : s(x)
    {
        if (!x || x->type != SEXPREC::REAL || !x->dim || x->dim->ints.size() != 2) throw std::runtime_error("not a numeric matrix");
    }
    int nrow() const { return s->dim->ints[0]; }
    int ncol() const { return s->dim->ints[1]; }
    double *begin() { return s->real.data(); }
};

template <typename T> T as(SEXP x);
template <> inline double as<double>(SEXP x) { if (!x || x->type != SEXPREC::REAL || x->real.size() != 1) throw std::runtime_error("expecting a single numeric value"); return x->real[0]; }
template <> inline int as<int>(SEXP x)
{
    if (x && x->type == SEXPREC::INT && x->ints.size() == 1) return x->ints[0];
    if (x && x->type == SEXPREC::REAL && x->real.size() == 1) return (int)x->real[0];
    throw std::runtime_error("expecting a single integer value");
}
template <> inline bool as<bool>(SEXP x) { if (!x || x->type != SEXPREC::LGL || x->ints.size() != 1) throw std::runtime_error("expecting a single logical value"); return x->ints[0] != 0; }

class List {
public:
    SEXP s;
    List() : s(rstub::alloc(SEXPREC::LIST)) {}
    template <typename... A> static List create(const A &...a) { List l; (l.push(a), ...); return l; }
    void push(const NumericVector &v) { s->items.push_back(v.s); }
    void push(int lgl) { SEXP x = rstub::alloc(SEXPREC::LGL); x->ints.push_back(lgl); s->items.push_back(x); }
    operator SEXP() const { return s; }
};

inline void stop(const char *fmt, ...)
{
    char buf[1024];
    va_list ap; va_start(ap, fmt); std::vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    throw std::runtime_error(buf);
}

} // namespace Rcpp
