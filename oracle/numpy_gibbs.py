"""Statistical oracle: an independent NumPy restatement of the reference's Gibbs sampler that uses
NumPy's own library samplers (Generator.multinomial / gamma / uniform) instead of the draw
specification shared by gibbs_oracle.c and the CUDA kernels.

TEST INFRASTRUCTURE ONLY.  Parity unpinned (no reference tests, R unavailable): this file exists so
that the custom samplers (Philox streams, inversion/BTRS binomials, direct method, Marsaglia-Tsang)
are checked DISTRIBUTIONALLY against trusted library code running the same model:
posterior summaries of the CUDA path must agree with this sampler within the tolerances of
BASELINE.json (Hungarian-matched signature cosine >= 0.99, median exposure relative error <= 5 %,
same BIC-optimal nsig).

Follows src/gibbs_2.cpp:85-251 (the seven conditional updates, clamps at 1e-160, the X==0 case of
the MH steps), :303-318 (fresh random permutation of the steps every sweep, Pfixed guards),
:320-334 (sample storage) and R/cpp_eBayesNMF.R:180-186 (log-likelihood, BIC).
"""
import numpy as np
from scipy.special import gammaln

TINY = 1e-160


def _gamma(rng, shape, rate):
    shape = np.maximum(TINY, shape)                       # one_gamma_dist, src/gibbs_2.cpp:39-43
    scale = np.maximum(TINY, 1.0 / rate)
    return rng.gamma(shape, scale)


def loglik(M, W, P, E):
    lam = (P @ E) * W
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(np.sum(M * np.log(lam) - gammaln(M + 1.0) - lam))


def gibbs_run(M, W, P, E, Ap, Bp, Ae, Be, hyper, burn, eval, Pfixed=False, seed=0):
    rng = np.random.default_rng(seed)
    M = np.trunc(np.asarray(M, dtype=np.float64)); W = np.asarray(W, dtype=np.float64)
    P, E, Ap, Bp, Ae, Be = (np.array(x, dtype=np.float64) for x in (P, E, Ap, Bp, Ae, Be))
    I, J = M.shape
    K = P.shape[1]
    ap, bp, ae, be, lp, le, var_ap, var_ae = hyper
    Mi = M.astype(np.int64).ravel()                       # cell (i,j) at i*J + j
    Ps = np.zeros((I, K, eval)); Es = np.zeros((K, J, eval)); ll = np.zeros(eval)

    def draw_z():
        pr = P[:, None, :] * E.T[None, :, :]              # I x J x K
        pr = (pr / pr.sum(axis=2, keepdims=True)).reshape(I * J, K)
        return rng.multinomial(Mi, pr).reshape(I, J, K)

    Z = draw_z()                                           # initial allocation, R/cpp_eBayesNMF.R:79-89

    def mh(A, X, B, l, var):
        Y = np.maximum(TINY, _gamma(rng, A * A / var, A / var))
        with np.errstate(all="ignore"):
            lnp = ((Y - A) * (np.log(B) + np.log(X) - l) + gammaln(A + 1) - gammaln(Y + 1)
                   + gammaln(A * A / var) - gammaln(Y * Y / var)
                   + ((Y * Y - A * A) / var) * np.log(A * Y / var) + np.log(Y) - np.log(A))
            pch = np.exp(lnp)
        pch = np.where(X == 0, (Y < A).astype(float), pch)
        U = rng.uniform(size=A.shape)
        acc = U <= pch                                     # NaN compares False -> reject
        return np.where(acc, Y, A)

    for k in range(burn + eval):
        for step in rng.permutation(7) + 1:
            if step == 1:
                Z = draw_z()
            elif step == 2 and not Pfixed:
                P = _gamma(rng, Ap + 1 + Z.sum(axis=1), Bp + W @ E.T)
            elif step == 3:
                E = _gamma(rng, Ae + 1 + Z.sum(axis=0).T, Be + P.T @ W)
            elif step == 4 and not Pfixed:
                Bp = np.maximum(TINY, _gamma(rng, Ap + ap + 1, P + bp))
            elif step == 5:
                Be = np.maximum(TINY, _gamma(rng, Ae + ae + 1, E + be))
            elif step == 6 and not Pfixed:
                Ap = mh(Ap, P, Bp, lp, var_ap)
            elif step == 7:
                Ae = mh(Ae, E, Be, le, var_ae)
        if k >= burn:
            r = k - burn
            Ps[:, :, r] = P; Es[:, :, r] = E
            ll[r] = loglik(M, W, P, E)
    bic = 2.0 * ll - K * (I + J) * np.log(J)
    return dict(Ps=Ps, Es=Es, loglik=ll, bic=bic, P=P, E=E, Ap=Ap, Bp=Bp, Ae=Ae, Be=Be)


def summaries(Ps, Es):
    """Normalised posterior medians (R/Definition_SignExp_object.R:116-122 and :259,:297)."""
    s = Ps.sum(axis=0)                                     # K x R
    Sign = Ps / s[None, :, :]
    Exp = Es * s[:, None, :]
    return np.median(Sign, axis=2), np.median(Exp, axis=2)


def match_signatures(Pa, Pb):
    """Hungarian matching on cosine similarity; returns (permutation of b's columns, cosines)."""
    from scipy.optimize import linear_sum_assignment
    A = Pa / np.linalg.norm(Pa, axis=0, keepdims=True)
    B = Pb / np.linalg.norm(Pb, axis=0, keepdims=True)
    C = A.T @ B
    r, c = linear_sum_assignment(-C)
    return c, C[r, c]
