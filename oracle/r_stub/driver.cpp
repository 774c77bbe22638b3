// Drives the reference-side binding r_shim/signer_gibbs_shim.cpp -- compiled unmodified against the stand-in Rcpp.h
// of this directory -- the way R's .Call would: builds the 24 SEXP arguments of _signeR_GibbsSamplerCpp
// (src/RcppExports.cpp:31, src/signeR_init.c:13) from raw arrays, calls it, and unpacks the returned list
// (src/gibbs_2.cpp:406-423) while checking its shape: 8 slots, the NA pattern of keep_par, the named dim attributes
// cube_to_rcpp gives the cubes (src/gibbs_2.cpp:72-82).  TEST INFRASTRUCTURE ONLY (tests/test_gpu_rshim.py).
#include "Rcpp.h"

extern "C" SEXP _signeR_GibbsSamplerCpp(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP,
                                        SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);

static SEXP real_array(const double *p, std::initializer_list<int> dims)
{
    SEXP x = rstub::alloc(SEXPREC::REAL);
    size_t n = 1;
    for (int d : dims) n *= (size_t)d;
    x->real.assign(p, p + n);
    if (dims.size() > 1) { x->dim = rstub::alloc(SEXPREC::INT); for (int d : dims) x->dim->ints.push_back(d); }
    return x;
}
static SEXP real1(double v) { return real_array(&v, {1}); }
static SEXP int1(int v) { SEXP x = rstub::alloc(SEXPREC::INT); x->ints.push_back(v); return x; }
static SEXP lgl1(int v) { SEXP x = rstub::alloc(SEXPREC::LGL); x->ints.push_back(v); return x; }

static bool dims_are(SEXP x, std::initializer_list<const char *> names, std::initializer_list<int> dims)
{
    if (!x->dim || x->dim->ints.size() != dims.size() || x->dim->names.size() != names.size()) return false;
    size_t t = 0;
    for (int d : dims) if (x->dim->ints[t++] != d) return false;
    t = 0;
    for (const char *n : names) if (x->dim->names[t++] != n) return false;
    return true;
}

// the two unif_rand() draws the shim turns into the 64-bit Philox key, for the test to replay
extern "C" uint64_t rshim_seed_for(uint64_t r_seed)
{
    rstub::rng_state() = r_seed;
    const uint64_t hi = (uint64_t)(rstub::unif() * 4294967296.0), lo = (uint64_t)(rstub::unif() * 4294967296.0);
    return (hi << 32) | lo;
}

// returns 0 ok, 1 R error (text in err), 2 malformed list.  Output pointers may be null.
extern "C" int rshim_call(int I, int J, int K, const double *M, const double *W, const double *Z, const double *P, const double *E,
                          const double *Ap, const double *Bp, const double *Ae, const double *Be, const double hyper[8], int burn, int eval,
                          int Pfixed, int Zfixed, int Thetafixed, int Afixed, int keep_par, uint64_t r_seed, int w_cols /* columns of the W argument; J unless a test wants a mismatch */,
                          double *Zs, double *Ps, double *Es, double *Aps, double *Bps, double *Aes, double *Bes, char *err, int err_len)
{
    rstub::release_all();
    rstub::rng_state() = r_seed;                               // set.seed(r_seed)
    SEXP a[24] = {real_array(M, {I, J}), real_array(W, {I, w_cols}), real_array(Z, {I, J, K}), real_array(P, {I, K}), real_array(E, {K, J}),
                  real_array(Ap, {I, K}), real_array(Bp, {I, K}), real_array(Ae, {K, J}), real_array(Be, {K, J}),
                  real1(hyper[0]), real1(hyper[1]), real1(hyper[2]), real1(hyper[3]), real1(hyper[4]), real1(hyper[5]), real1(hyper[6]),
                  real1(hyper[7]), int1(burn), int1(eval), lgl1(Pfixed), lgl1(Zfixed), lgl1(Thetafixed), lgl1(Afixed), lgl1(keep_par)};
    SEXP r = _signeR_GibbsSamplerCpp(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13], a[14], a[15], a[16],
                                     a[17], a[18], a[19], a[20], a[21], a[22], a[23]);
    int rc = 0;
    if (!r || r->type == SEXPREC::ERROR) {
        if (err && err_len > 0) { std::snprintf(err, (size_t)err_len, "%s", r ? r->msg.c_str() : "null result"); }
        rc = 1;
    } else {
        auto is_na = [](SEXP x) { return x->type == SEXPREC::LGL && x->ints.size() == 1 && x->ints[0] == NA_LOGICAL; };
        auto cube = [](SEXP x) { return x->type == SEXPREC::REAL; };
        bool ok = r->type == SEXPREC::LIST && r->items.size() == 8 && is_na(r->items[1]);
        if (ok) {
            SEXP *it = r->items.data();
            ok = cube(it[2]) && cube(it[3]) && cube(it[5]) && cube(it[7]) && dims_are(it[2], {"i", "n", "r"}, {I, K, eval}) &&
                 dims_are(it[3], {"n", "j", "r"}, {K, J, eval}) && dims_are(it[5], {"i", "n", "r"}, {I, K, eval}) &&
                 dims_are(it[7], {"n", "j", "r"}, {K, J, eval});
            if (keep_par)
                ok = ok && cube(it[0]) && cube(it[4]) && cube(it[6]) && dims_are(it[0], {"i", "j", "n", "r"}, {I, J, K, 1}) &&
                     dims_are(it[4], {"i", "n", "r"}, {I, K, eval}) && dims_are(it[6], {"n", "j", "r"}, {K, J, eval});
            else
                ok = ok && is_na(it[0]) && is_na(it[4]) && is_na(it[6]);
            if (ok) {
                auto cp = [](double *dst, SEXP x) { if (dst && x->type == SEXPREC::REAL) std::memcpy(dst, x->real.data(), sizeof(double) * x->real.size()); };
                cp(Zs, it[0]); cp(Ps, it[2]); cp(Es, it[3]); cp(Aps, it[4]); cp(Bps, it[5]); cp(Aes, it[6]); cp(Bes, it[7]);
            }
        }
        if (!ok) rc = 2;
    }
    rstub::release_all();
    return rc;
}
