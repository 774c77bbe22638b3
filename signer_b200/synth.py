"""Synthetic cohorts and initial states for tests and benchmarks.

Generators, seeds and distributions follow the measurement contract: sparse COSMIC-like signatures,
log-normal mutation burdens, opportunity W[i,j] = w_i * c_j, counts M ~ Poisson((P E) o W).
"""
import os

import numpy as np

HYPER_DEFAULT = dict(ap=0.8005, bp=0.043, ae=2.0506, be=1.325, lp=5.3329, le=1.436)   # R/signeR.R:110
VAR_DEFAULT = dict(var_ap=10.0, var_ae=10.0)                                             # R/signeR.R:65

_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "breast21.npz")

# BASELINE.json configs: name -> (I, J, K_true, burden median)
CONFIGS = {
    "cfg2_96x560": dict(I=96, J=560, K_true=5, burden=2000.0, seed=20240502),
    "cfg3_96x10000_K20": dict(I=96, J=10000, K_true=20, burden=2000.0, seed=20240503),
    "cfg4_1536x20000_K30": dict(I=1536, J=20000, K_true=30, burden=20000.0, seed=20240504),
    "cfg5_96x100000": dict(I=96, J=100000, K_true=12, burden=2000.0, seed=20240505),
}


def breast21():
    """The reference's vignette data, contexts x genomes (tests/golden/make_fixtures.py)."""
    d = np.load(_GOLDEN)
    return np.asfortranarray(d["M"]), np.asfortranarray(d["W"]), np.asfortranarray(d["cosmic"])


def cohort(I, J, K_true, burden=2000.0, seed=0, opportunity=True):
    """Synthetic (M, W, P_true, E_true), column-major float64."""
    rng = np.random.Generator(np.random.Philox(seed))
    P_true = rng.dirichlet(np.full(I, 0.1), size=K_true).T                 # I x K, columns sum to 1
    N = rng.lognormal(np.log(burden), 1.0, size=J)
    E_true = rng.dirichlet(np.full(K_true, 0.5), size=J).T * N            # K x J
    if opportunity:
        if I % 96 == 0 and os.path.exists(_GOLDEN):
            w0 = np.load(_GOLDEN)["W"][:, 0]
            w = np.tile(w0 / w0.mean(), I // 96)
        else:
            w = rng.lognormal(0.0, 0.5, size=I)
            w = w / w.mean()
        c = rng.uniform(0.8, 1.2, size=J)
        W = np.outer(w, c)
    else:
        W = np.ones((I, J))
    M = rng.poisson((P_true @ E_true) * W).astype(np.float64)
    return np.asfortranarray(M), np.asfortranarray(W), np.asfortranarray(P_true), np.asfortranarray(E_true)


def named(name):
    c = CONFIGS[name]
    return cohort(c["I"], c["J"], c["K_true"], c["burden"], c["seed"])


def init_state(I, J, K, hyper=None, seed=0, fixedP=None):
    """start='random' initial state (R/cpp_eBayesNMF.R:35-51): Ap~Exp(lp), Bp~Gamma(ap,bp),
    Ae~Exp(le), Be~Gamma(ae,be), P~Gamma(Ap+1,Bp), E~Gamma(Ae+1,Be) (rates)."""
    h = dict(HYPER_DEFAULT)
    h.update(hyper or {})
    rng = np.random.Generator(np.random.Philox(seed))
    if fixedP is not None:
        P = np.asfortranarray(np.asarray(fixedP, dtype=np.float64))
        Ap = np.zeros((I, K), order="F"); Bp = np.zeros((I, K), order="F")        # R/cpp_eBayesNMF.R:36-38
    else:
        Ap = rng.exponential(1.0 / h["lp"], size=(I, K))
        Bp = rng.gamma(h["ap"], 1.0 / h["bp"], size=(I, K))
        P = rng.gamma(Ap + 1.0, 1.0 / Bp)
    Ae = rng.exponential(1.0 / h["le"], size=(K, J))
    Be = rng.gamma(h["ae"], 1.0 / h["be"], size=(K, J))
    E = rng.gamma(Ae + 1.0, 1.0 / Be)
    return {k: np.asfortranarray(v) for k, v in dict(P=P, E=E, Ap=Ap, Bp=Bp, Ae=Ae, Be=Be).items()}


def hyper_tuple(hyper=None):
    h = dict(HYPER_DEFAULT); h.update(VAR_DEFAULT); h.update(hyper or {})
    return tuple(h[k] for k in ("ap", "bp", "ae", "be", "lp", "le", "var_ap", "var_ae"))
