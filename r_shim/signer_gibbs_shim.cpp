// r_shim/signer_gibbs_shim.cpp -- the reference-side binding: replaces src/gibbs_2.cpp AND the
// GibbsSamplerCpp half of src/RcppExports.cpp in TojalLab/signeR (v2.3.3).
//
// It exports the SAME .Call symbol with the SAME 24 SEXP arguments the package registers
// (src/signeR_init.c:13  {"_signeR_GibbsSamplerCpp", ..., 24};  R/RcppExports.R:8-10), so
// R/cpp_eBayesNMF.R:104-107,165-168 and R/Definition_SignExp_object.R:438,506,571,633,683 call it
// unchanged, and returns the same positional list (src/gibbs_2.cpp:406-423) with the same named
// dim attributes (cube_to_rcpp, src/gibbs_2.cpp:72-82) -- R code indexes Ps[,,'r'=r].
//
// R, Rcpp and RcppArmadillo are absent from this image and from the GPU box, so the file cannot be built into the
// package here.  It IS compiled, unmodified, against a stand-in Rcpp.h (oracle/r_stub, test infrastructure) and its
// entry point is executed on the GPU by tests/test_gpu_rshim.py, which compares the returned list with the one the
// reference's own src/gibbs_2.cpp returns on the same inputs (oracle/_ref).
//
// Build inside the package: r_shim/Makevars replaces src/Makevars (src/fuzzy.cpp and the registration in
// src/signeR_init.c stay as they are).
#include <Rcpp.h>
#include "signer_gibbs.h"

using namespace Rcpp;

static NumericVector named_cube(int d0, int d1, int d2, const char *n0, const char *n1, const char *n2)
{
    NumericVector v((R_xlen_t)d0 * d1 * d2);
    IntegerVector dims = IntegerVector::create(_[n0] = d0, _[n1] = d1, _[n2] = d2);
    v.attr("dim") = dims;
    return v;
}

RcppExport SEXP _signeR_GibbsSamplerCpp(SEXP MSEXP, SEXP WSEXP, SEXP ZSEXP, SEXP PSEXP, SEXP ESEXP, SEXP ApSEXP, SEXP BpSEXP,
                                        SEXP AeSEXP, SEXP BeSEXP, SEXP apSEXP, SEXP bpSEXP, SEXP aeSEXP, SEXP beSEXP, SEXP lpSEXP,
                                        SEXP leSEXP, SEXP var_apSEXP, SEXP var_aeSEXP, SEXP burnSEXP, SEXP evalSEXP,
                                        SEXP PfixedSEXP, SEXP ZfixedSEXP, SEXP ThetafixedSEXP, SEXP AfixedSEXP, SEXP keep_parSEXP)
{
BEGIN_RCPP
    RNGScope scope;                                   // src/RcppExports.cpp:34: set.seed() still controls the run
    NumericMatrix M(MSEXP), W(WSEXP), P(PSEXP), E(ESEXP), Ap(ApSEXP), Bp(BpSEXP), Ae(AeSEXP), Be(BeSEXP);
    NumericVector Z(ZSEXP);                           // i x j x n array
    const int I = M.nrow(), J = M.ncol(), K = P.ncol();
    if (W.nrow() != I || W.ncol() != J || E.nrow() != K || E.ncol() != J || P.nrow() != I ||
        Z.size() != (R_xlen_t)I * J * K)
        stop("GibbsSamplerCpp: dimensions of the arguments do not match");

    signer_problem pr = { I, J, K, M.begin(), W.begin() };
    signer_state st = { Z.begin(), P.begin(), E.begin(), Ap.begin(), Bp.begin(), Ae.begin(), Be.begin() };
    signer_opts op;
    op.ap = as<double>(apSEXP); op.bp = as<double>(bpSEXP); op.ae = as<double>(aeSEXP); op.be = as<double>(beSEXP);
    op.lp = as<double>(lpSEXP); op.le = as<double>(leSEXP);
    op.var_ap = as<double>(var_apSEXP); op.var_ae = as<double>(var_aeSEXP);
    op.burn = as<int>(burnSEXP); op.eval = as<int>(evalSEXP);
    op.Pfixed = as<bool>(PfixedSEXP); op.Zfixed = as<bool>(ZfixedSEXP); op.Thetafixed = as<bool>(ThetafixedSEXP);
    op.Afixed = as<bool>(AfixedSEXP); op.keep_par = as<bool>(keep_parSEXP);
    // a 64-bit Philox key from R's generator: two unif_rand() draws of 32 bits each
    const uint64_t hi = (uint64_t)(unif_rand() * 4294967296.0), lo = (uint64_t)(unif_rand() * 4294967296.0);
    op.seed = (hi << 32) | lo;
    op.sweep0 = 0;
    op.device = 0;                                    // or getOption("signeR.device") / worker index mod #GPUs
    op.forced_order = NULL;

    const int R = op.eval;
    const bool keep = op.keep_par;
    NumericVector Ps = named_cube(I, K, R, "i", "n", "r"), Es = named_cube(K, J, R, "n", "j", "r");
    NumericVector Bps = named_cube(I, K, R, "i", "n", "r"), Bes = named_cube(K, J, R, "n", "j", "r");
    NumericVector Aps, Aes, Zs;
    signer_outputs out;
    memset(&out, 0, sizeof(out));
    out.Ps = Ps.begin(); out.Es = Es.begin(); out.Bps = Bps.begin(); out.Bes = Bes.begin();
    if (keep) {
        Aps = named_cube(I, K, R, "i", "n", "r"); Aes = named_cube(K, J, R, "n", "j", "r");
        Zs = NumericVector((R_xlen_t)I * J * K);
        Zs.attr("dim") = IntegerVector::create(_["i"] = I, _["j"] = J, _["n"] = K, _["r"] = 1);   // src/gibbs_2.cpp:279-282
        out.Aps = Aps.begin(); out.Aes = Aes.begin(); out.Zs = Zs.begin();
    }
    const int rc = signer_gibbs_run(&pr, &st, &op, &out);
    if (rc != SIGNER_OK) stop("GibbsSamplerCpp (CUDA): %s (status %d)", signer_last_error(), rc);
    if (keep)                                          // src/gibbs_2.cpp:406-413
        return List::create(Zs, NA_LOGICAL, Ps, Es, Aps, Bps, Aes, Bes);
    return List::create(NA_LOGICAL, NA_LOGICAL, Ps, Es, NA_LOGICAL, Bps, NA_LOGICAL, Bes);       // :415-423
END_RCPP
}
