"""eBayesNMF: the per-nsig driver, mirroring R/cpp_eBayesNMF.R:30-211 around the CUDA sampler.

Same arguments, same order of operations, same positional result list.  Differences, all on the
host side of the boundary and all stated in DESIGN.md:
  * random draws for the initial state come from numpy (R's RNG is not available);
  * the initial Z (the R loop :79-89, I*J rmultinom calls) is drawn on the device (Z=None);
  * the EM loop (:97-157) keeps the chain resident on the device between iterations and feeds the
    M-step from three device-reduced sufficient statistics instead of the Aps/Bps/Aes/Bes cubes;
  * the per-sample log-likelihood / BIC (:180-186) comes from the device;
  * start != 'random' uses a small in-house multiplicative-update NMF instead of NMF::nmf.
"""
import sys

import numpy as np

from . import gibbs
from .estimate import estimate_from_stats
from .signexp import SignExp


def _nmf_init(V, n, method, rng, iters=300):
    """Stand-in for NMF::nmf(Mcor/Wcor, n, method=start) (R/cpp_eBayesNMF.R:57): basis and coef
    from multiplicative updates (Lee & Seung).  'lee'/'Frobenius'-like methods use the Euclidean
    updates, 'brunet'/'KL'-like methods the KL updates."""
    V = np.maximum(np.asarray(V, dtype=np.float64), 0.0)
    i, j = V.shape
    scale = np.sqrt(max(V.mean(), 1e-12) / n)
    Wm = rng.uniform(0.1, 1.0, size=(i, n)) * scale
    H = rng.uniform(0.1, 1.0, size=(n, j)) * scale
    kl = str(method).lower() in ("brunet", "kl", "nsnmf", "sinmf")
    eps = 1e-12
    for _ in range(iters):
        if kl:
            WH = Wm @ H + eps
            H *= (Wm.T @ (V / WH)) / (Wm.sum(axis=0)[:, None] + eps)
            WH = Wm @ H + eps
            Wm *= ((V / WH) @ H.T) / (H.sum(axis=1)[None, :] + eps)
        else:
            H *= (Wm.T @ V) / (Wm.T @ Wm @ H + eps)
            Wm *= (V @ H.T) / (Wm @ (H @ H.T) + eps)
    return np.maximum(Wm, 1e-300), np.maximum(H, 1e-300)


def _mle_exposures(M, W, P, iters=500):
    """E maximising the Poisson likelihood of M given fixed P (R/cpp_eBayesNMF.R:63-77, per-genome
    nloptr SBPLX there): EM / KL multiplicative updates converge to the same maximiser."""
    n = P.shape[1]
    E = np.ones((n, M.shape[1]))
    eps = 1e-300
    for _ in range(iters):
        lam = (P @ E) * W + eps
        E *= (P.T @ (W * (M / lam))) / (P.T @ W + eps)
    return E


def initial_state(M, W, n, ap, bp, ae, be, lp, le, P=None, fixP=False, start="random", rng=None):
    """The chain's starting point exactly as R/cpp_eBayesNMF.R:35-77 builds it (numpy draws instead
    of R's).  Returns P, E, Ap, Bp, Ae, Be."""
    rng = rng if rng is not None else np.random.default_rng(gibbs.next_seed())
    i, j = M.shape
    if fixP:
        P = np.asfortranarray(np.asarray(P, dtype=np.float64))
        n = P.shape[1]
        Ap = np.zeros((i, n)); Bp = np.zeros((i, n))                      # :36-38
    else:
        Ap = rng.exponential(1.0 / lp, size=(i, n))                        # :40-41
        Bp = rng.gamma(ap, 1.0 / bp, size=(i, n))
    Ae = rng.exponential(1.0 / le, size=(n, j))                            # :43-44
    Be = rng.gamma(ae, 1.0 / be, size=(n, j))
    if P is None:
        if start == "random":                                              # :46-51
            P = rng.gamma(Ap + 1.0, 1.0 / Bp)
            E = rng.gamma(Ae + 1.0, 1.0 / Be)
        else:                                                              # :52-62
            Mcor = M.copy()
            if min(Mcor.sum(axis=0).min(), Mcor.sum(axis=1).min()) == 0:
                Mcor[Mcor == 0] = 0.01
            Wcor = W.copy()
            Wcor[Wcor == 0] = 0.01
            P, E = _nmf_init(Mcor / Wcor, n, start, rng)
    else:
        P = np.asfortranarray(np.asarray(P, dtype=np.float64))
        E = _mle_exposures(M, W, P)                                        # :63-77
    return tuple(np.asfortranarray(x) for x in (P, E, Ap, Bp, Ae, Be))


def eBayesNMF(M, W, n, ap, bp, ae, be, lp, le, var_ap=10, var_ae=10, P=None, fixP=False, start="random",
              burn_it=500, eval_it=1000, estimate_hyper=False, EM_lim=100, EM_eval_it=100, keep_param=False,
              *, rng=None, device=0, verbose=False, keep_samples=True, device_summaries=False):
    """Returns the reference's list: [Phat, Ehat, Bics, SE, Aps, Bps, Aes, Bes, hyper] with the four
    parameter cubes None unless keep_param (R's list positions [[1]]..[[9]]).
    keep_samples=False, device_summaries=True: Phat/Ehat come from the device (normalisation and medians
    computed there) and no sample cube is copied to the host; SE is then None."""
    rng = rng if rng is not None else np.random.default_rng(gibbs.next_seed())
    M = np.asfortranarray(np.asarray(M, dtype=np.float64))
    W = np.asfortranarray(np.asarray(W, dtype=np.float64))
    i, j = M.shape
    P, E, Ap, Bp, Ae, Be = initial_state(M, W, n, ap, bp, ae, be, lp, le, P, fixP, start, rng)
    n = P.shape[1]
    apvet, bpvet, lpvet = ([np.nan], [np.nan], [np.nan]) if fixP else ([ap], [bp], [lp])   # :91-96
    aevet, bevet, levet = [ae], [be], [le]

    # the chain lives on the device from here on; the initial Z is drawn there (:79-89)
    chain = gibbs.Chain(M, W, P, E, Ap, Bp, Ae, Be, (ap, bp, ae, be, lp, le, var_ap, var_ae), Z=None, Pfixed=fixP,
                        seed=int(rng.integers(0, 2 ** 63 - 1)), device=device)
    try:
        if estimate_hyper:                                                 # :97-157
            burn_HO, upgrade, it = burn_it, 100.0, 0
            if verbose:
                print("EM algorithm:", file=sys.stderr)
            while upgrade > 0.05 and it < EM_lim:
                chain.run(burn=burn_HO, eval=EM_eval_it, keep_par=True, fetch=False)
                st = chain.stats()
                aeold, beold, leold = ae, be, le
                ae, be, le = estimate_from_stats(st[4:8], aeold)           # :124-127
                aevet.append(ae); bevet.append(be); levet.append(le)
                if fixP:
                    upgrade = max(abs(ae - aeold), abs(be - beold), abs(le - leold))
                else:
                    apold, bpold, lpold = ap, bp, lp
                    ap, bp, lp = estimate_from_stats(st[0:4], apold)       # :141-144
                    apvet.append(ap); bpvet.append(bp); lpvet.append(lp)
                    upgrade = max(abs(ap - apold), abs(bp - bpold), abs(ae - aeold), abs(be - beold),
                                  abs(lp - lpold), abs(le - leold))
                chain.set_hyper((ap, bp, ae, be, lp, le, var_ap, var_ae))
                it += 1
                burn_HO = 200                                               # :152
        if verbose:
            print("Running  Gibbs sampler for %d signature%s..." % (n, "" if n == 1 else "s"), file=sys.stderr)
        res = chain.run(burn=burn_it, eval=eval_it, keep_par=keep_param, fetch=True, want_Z=False,
                        samples=keep_samples, summaries=(device_summaries and not keep_samples))   # :165-168
    finally:
        chain.close()
    Bics = np.array(res.bic)                                               # :180-186, computed on the device
    out = [None, None, Bics, None, None, None, None, None, None]
    if keep_samples:
        Ps, Es = res[2], res[3]
        SE = SignExp(Ps, Es)                                               # :187-189
        Phat, Ehat = SE.median_sign(), SE.median_exp()
        if not fixP:                                                       # :190-198 order by total exposure
            totalexp = Ehat.sum(axis=1) * Phat.sum(axis=0)
            order = np.argsort(-totalexp, kind="stable")
            Ps = np.asfortranarray(Ps[:, order, :]); Es = np.asfortranarray(Es[order, :, :])
            SE = SignExp(Ps, Es)
            Phat, Ehat = SE.median_sign(), SE.median_exp()
        out[0], out[1], out[3] = Phat, Ehat, SE
        if keep_param:
            out[4], out[5], out[6], out[7] = res[4], res[5], res[6], res[7]
    elif device_summaries:                                                 # :187-198 on the device, re-ordering here
        Phat, Ehat = res.Phat, res.Ehat
        if not fixP:
            order = np.argsort(-(Ehat.sum(axis=1) * Phat.sum(axis=0)), kind="stable")
            Phat, Ehat = np.asfortranarray(Phat[:, order]), np.asfortranarray(Ehat[order, :])
        out[0], out[1] = Phat, Ehat
    if estimate_hyper:                                                     # :202-208
        out[8] = dict(ap=np.array(apvet), bp=np.array(bpvet), ae=np.array(aevet), be=np.array(bevet),
                      lp=np.array(lpvet), le=np.array(levet))
    else:
        out[8] = dict(ap=np.array([ap]), bp=np.array([bp]), ae=np.array([ae]), be=np.array([be]),
                      lp=np.array([lp]), le=np.array([le]))
    return out
