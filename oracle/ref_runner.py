"""Runs the reference's own src/gibbs_2.cpp (compiled by oracle/ref_build into oracle/_ref/) with the
third-party primitives replaced by the draw specification.  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libref_gibbs.so")
_LIB = None


def available():
    return os.path.exists(SO)


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(SO)
        dp = C.POINTER(C.c_double)
        L.ref_gibbs_run.restype = C.c_int
        L.ref_gibbs_run.argtypes = ([C.c_int] * 3 + [dp] * 9 + [dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint32]
                                    + [dp] * 7 + [dp, C.POINTER(C.c_int)])
        _LIB = L
    return _LIB


def _f(a):
    return np.asfortranarray(np.array(a, dtype=np.float64, copy=True))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def gibbs_run(M, W, Z, P, E, Ap, Bp, Ae, Be, hyper=None, burn=0, eval=1, Pfixed=False, keep_par=True, seed=1, sweep0=0,
              want_Z=True, **_):
    """Same calling convention as oracle.gibbs_run.  The reference takes Z by value and needs one:
    when Z is None the oracle's initial allocation (stream 8) is used, like the GPU path does."""
    from . import oracle as O
    h = dict(O.HYPER_DEFAULT); h.update(hyper or {})
    hv = np.array([h[k] for k in ("ap", "bp", "ae", "be", "lp", "le", "var_ap", "var_ae")], dtype=np.float64)
    M, W = _f(M), _f(W)
    I, J = M.shape
    P, E, Ap, Bp, Ae, Be = (_f(x) for x in (P, E, Ap, Bp, Ae, Be))
    K = P.shape[1]
    if Z is None:
        Z = O.gibbs_run(M, W, None, P, E, Ap, Bp, Ae, Be, hyper=hyper, burn=0, eval=0, seed=seed, sweep0=sweep0)["Z"]
    Zc = _f(Z)
    R = eval
    f = dict(order="F")
    Ps = np.zeros((I, K, R), **f); Es = np.zeros((K, J, R), **f); Aps = np.zeros((I, K, R), **f); Bps = np.zeros((I, K, R), **f)
    Aes = np.zeros((K, J, R), **f); Bes = np.zeros((K, J, R), **f); Zs = np.zeros((I, J, K), **f)
    dev = C.c_double(0.0); ok = C.c_int(0)
    rc = lib().ref_gibbs_run(I, J, K, _p(M), _p(W), _p(Zc), _p(P), _p(E), _p(Ap), _p(Bp), _p(Ae), _p(Be), _p(hv), burn, eval,
                             int(Pfixed), int(keep_par), seed, sweep0, _p(Ps), _p(Es), _p(Aps), _p(Bps), _p(Aes), _p(Bes), _p(Zs),
                             C.byref(dev), C.byref(ok))
    if rc != 0:
        raise RuntimeError("ref_gibbs_run failed: %d" % rc)
    return dict(Ps=Ps, Es=Es, Aps=Aps, Bps=Bps, Aes=Aes, Bes=Bes, Z=Zs, max_prob_dev=dev.value, list_shape_ok=bool(ok.value))
