// Runner around the reference's own GibbsSamplerCpp (src/gibbs_2.cpp, compiled from /root/reference
// against the stand-in headers in shim/).  TEST INFRASTRUCTURE ONLY.
#include "RcppArmadillo.h"
#include <cstring>

Rcpp::List GibbsSamplerCpp(arma::mat M, arma::mat W, arma::cube Z, arma::mat P, arma::mat E, arma::mat Ap, arma::mat Bp,
                           arma::mat Ae, arma::mat Be, double ap, double bp, double ae, double be, double lp, double le,
                           double var_ap, double var_ae, int burn, int eval, bool Pfixed, bool Zfixed, bool Thetafixed,
                           bool Afixed, bool keep_par);

static arma::mat mk(const double *p, int r, int c) { arma::mat m(r, c); std::memcpy(m.d.data(), p, sizeof(double) * (size_t)r * c); return m; }

extern "C" int ref_gibbs_run(int I, int J, int K, const double *M, const double *W, const double *Z, const double *P,
                             const double *E, const double *Ap, const double *Bp, const double *Ae, const double *Be,
                             const double hyper[8], int burn, int eval, int Pfixed, int keep_par, uint64_t seed, uint32_t sweep0,
                             double *Ps, double *Es, double *Aps, double *Bps, double *Aes, double *Bes, double *Zs,
                             double *max_prob_dev, int *dims_ok)
{
    try {
        refctx::Ctx &x = refctx::ctx();
        x = refctx::Ctx();
        x.seed = seed; x.sweep0 = sweep0; x.I = I; x.J = J; x.K = K; x.Pfixed = Pfixed != 0;
        x.P.assign(P, P + (size_t)I * K); x.E.assign(E, E + (size_t)K * J);
        arma::cube Zc(I, J, K);
        std::memcpy(Zc.d.data(), Z, sizeof(double) * (size_t)I * J * K);
        Rcpp::List out = GibbsSamplerCpp(mk(M, I, J), mk(W, I, J), Zc, mk(P, I, K), mk(E, K, J), mk(Ap, I, K), mk(Bp, I, K),
                                         mk(Ae, K, J), mk(Be, K, J), hyper[0], hyper[1], hyper[2], hyper[3], hyper[4], hyper[5],
                                         hyper[6], hyper[7], burn, eval, Pfixed != 0, false, false, false, keep_par != 0);
        if (out.items.size() != 8) return 10;
        auto cp = [](double *dst, const Rcpp::NumericVector &v) { if (dst && !v.is_na) std::memcpy(dst, v.d.data(), sizeof(double) * v.d.size()); };
        // positional list of src/gibbs_2.cpp:406-423
        cp(Zs, out.items[0]); cp(Ps, out.items[2]); cp(Es, out.items[3]); cp(Aps, out.items[4]); cp(Bps, out.items[5]);
        cp(Aes, out.items[6]); cp(Bes, out.items[7]);
        int ok = out.items[1].is_na ? 1 : 0;
        if (keep_par) ok &= !out.items[0].is_na && !out.items[4].is_na && !out.items[6].is_na;
        else ok &= out.items[0].is_na && out.items[4].is_na && out.items[6].is_na;
        const Rcpp::IntegerVector &dp = out.items[2].dim, &de = out.items[3].dim;
        ok &= dp.size() == 3 && dp.names[0] == "i" && dp.names[1] == "n" && dp.names[2] == "r" && dp[0] == I && dp[1] == K && dp[2] == eval;
        ok &= de.size() == 3 && de.names[0] == "n" && de.names[1] == "j" && de.names[2] == "r" && de[0] == K && de[1] == J && de[2] == eval;
        if (dims_ok) *dims_ok = ok;
        if (max_prob_dev) *max_prob_dev = x.max_prob_dev;
        return 0;
    } catch (const std::exception &e) {
        return 11;
    }
}
