"""ctypes front-end of the CPU oracle (oracle/gibbs_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  The product package (signer_b200/) never imports this module.

Parity unpinned: see the header of gibbs_oracle.c.  Arrays follow the reference's layouts
(column-major doubles: M,W I x J; P,Ap,Bp I x K; E,Ae,Be K x J; Z cube I x J x K;
sample cubes (i,n,r) / (n,j,r)), i.e. numpy arrays in Fortran order.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

HYPER_DEFAULT = dict(ap=0.8005, bp=0.043, ae=2.0506, be=1.325, lp=5.3329, le=1.436,
                     var_ap=10.0, var_ae=10.0)          # R/signeR.R:65,110


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "gibbs_oracle.c")
    tab = os.path.join(_HERE, "det_tables.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(tab)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        d, u32, u64, i32 = C.c_double, C.c_uint32, C.c_uint64, C.c_int
        dp = C.POINTER(C.c_double)
        for f in ("oracle_log", "oracle_exp", "oracle_lgamma"):
            getattr(L, f).restype = d
            getattr(L, f).argtypes = [d]
        L.oracle_normal.restype = d
        L.oracle_normal.argtypes = [u64, u32, u32, u32, u32]
        L.oracle_u53.restype = d
        L.oracle_u53.argtypes = [u32, u32]
        L.oracle_binomial.restype = i32
        L.oracle_binomial.argtypes = [u64, u32, u32, u32, i32, i32, d]
        L.oracle_gamma.restype = d
        L.oracle_gamma.argtypes = [u64, u32, u32, u32, d, d]
        L.oracle_perm.restype = None
        L.oracle_perm.argtypes = [u64, u32, C.POINTER(i32)]
        L.oracle_loglik.restype = d
        L.oracle_loglik.argtypes = [i32, i32, i32, dp, dp, dp, dp]
        L.oracle_multinomial_cell.restype = None
        L.oracle_multinomial_cell.argtypes = [u64, u32, u32, i32, i32, i32, i32, d, dp, dp, dp]
        L.oracle_multinomial_cell_literal.restype = None
        L.oracle_multinomial_cell_literal.argtypes = [u64, u32, u32, i32, i32, i32, i32, d, dp, dp, dp]
        L.oracle_gibbs_run.restype = i32
        L.oracle_gibbs_run.argtypes = ([i32, i32, i32, dp, dp, dp] + [dp] * 6 + [dp, i32, i32, i32, i32,
                                       u64, u32, i32, C.POINTER(i32)] + [dp] * 10)
        _LIB = L
    return _LIB


def _f(a):
    return np.asfortranarray(np.array(a, dtype=np.float64, copy=True))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def philox(ctr, key):
    cc = (C.c_uint32 * 4)(*ctr)
    kk = (C.c_uint32 * 2)(*key)
    out = (C.c_uint32 * 4)()
    lib().oracle_philox(cc, kk, out)
    return [int(x) for x in out]


def perm(seed, sweep):
    out = (C.c_int * 7)()
    lib().oracle_perm(seed, sweep, out)
    return [int(x) for x in out]


def loglik(M, W, P, E):
    M, W, P, E = _f(M), _f(W), _f(P), _f(E)
    I, J = M.shape
    K = P.shape[1]
    return lib().oracle_loglik(I, J, K, _p(M), _p(W), _p(P), _p(E))


def multinomial_cell(seed, sweep, stream, s, g, Mval, P, E):
    P, E = _f(P), _f(E)
    I, K = P.shape
    z = np.zeros(K)
    lib().oracle_multinomial_cell(seed, sweep, stream, I, K, s, g, float(Mval), _p(P), _p(E), _p(z))
    return z


def multinomial_cell_literal(seed, sweep, stream, s, g, Mval, P, E):
    """R's rmultinom structure literally (k ascending, no early stop) -- see oracle_multinomial_cell_literal."""
    P, E = _f(P), _f(E)
    I, K = P.shape
    z = np.zeros(K)
    lib().oracle_multinomial_cell_literal(seed, sweep, stream, I, K, s, g, float(Mval), _p(P), _p(E), _p(z))
    return z


def gibbs_run(M, W, Z, P, E, Ap, Bp, Ae, Be, hyper=None, burn=0, eval=1, Pfixed=False, keep_par=True,
              seed=1, sweep0=0, nshards=1, forced_order=None, want_Z=True):
    """Run the oracle chain.  Returns a dict with the final state, the sample cubes in the
    reference's shapes, per-sample log-likelihoods, BICs (R/cpp_eBayesNMF.R:186) and Zin/Znj."""
    h = dict(HYPER_DEFAULT)
    h.update(hyper or {})
    hv = np.array([h[k] for k in ("ap", "bp", "ae", "be", "lp", "le", "var_ap", "var_ae")], dtype=np.float64)
    M, W = _f(M), _f(W)
    I, J = M.shape
    P, E, Ap, Bp, Ae, Be = (_f(x) for x in (P, E, Ap, Bp, Ae, Be))
    K = P.shape[1]
    Zc = None if Z is None else _f(Z)
    R = eval
    Ps = np.zeros((I, K, R), order="F"); Es = np.zeros((K, J, R), order="F")
    Aps = np.zeros((I, K, R), order="F"); Aes = np.zeros((K, J, R), order="F")
    Bps = np.zeros((I, K, R), order="F"); Bes = np.zeros((K, J, R), order="F")
    Zl = np.zeros((I, J, K), order="F") if want_Z else None
    ll = np.zeros(R)
    Zin = np.zeros((I, K), order="F"); Znj = np.zeros((K, J), order="F")
    fo = None
    if forced_order is not None:
        fo_arr = np.ascontiguousarray(np.array(forced_order, dtype=np.int32).reshape(burn + eval, 7))
        fo = fo_arr.ctypes.data_as(C.POINTER(C.c_int))
    rc = lib().oracle_gibbs_run(I, J, K, _p(M), _p(W), _p(Zc), _p(P), _p(E), _p(Ap), _p(Bp), _p(Ae), _p(Be),
                                _p(hv), burn, eval, int(Pfixed), int(keep_par), seed, sweep0, nshards, fo,
                                _p(Ps), _p(Es), _p(Aps), _p(Bps), _p(Aes), _p(Bes), _p(Zl), _p(ll), _p(Zin), _p(Znj))
    if rc != 0:
        raise RuntimeError("oracle_gibbs_run failed: %d" % rc)
    bic = 2.0 * ll - K * (I + J) * np.log(J)
    return dict(P=P, E=E, Ap=Ap, Bp=Bp, Ae=Ae, Be=Be, Ps=Ps, Es=Es, Aps=Aps, Bps=Bps, Aes=Aes, Bes=Bes,
                Z=Zl, loglik=ll, bic=bic, Zin=Zin, Znj=Znj)
