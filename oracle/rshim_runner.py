"""Calls the reference-side binding r_shim/signer_gibbs_shim.cpp (compiled unmodified against oracle/r_stub/Rcpp.h into
oracle/_ref/librshim.so) through its 24-SEXP entry point, the way R's .Call would.  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "librshim.so")
_LIB = None


def build():
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "r_stub"), "-s"])
    return SO


def available():
    return os.path.exists(SO)


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(SO)
        dp = C.POINTER(C.c_double)
        L.rshim_call.restype = C.c_int
        L.rshim_call.argtypes = [C.c_int] * 3 + [dp] * 9 + [dp] + [C.c_int] * 7 + [C.c_uint64, C.c_int] + [dp] * 7 + [C.c_char_p, C.c_int]
        L.rshim_seed_for.restype = C.c_uint64
        L.rshim_seed_for.argtypes = [C.c_uint64]
        _LIB = L
    return _LIB


class RError(RuntimeError):
    """What R would raise (Rcpp::stop / an exception crossing END_RCPP)."""


def _f(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def seed_for(r_seed):
    """The 64-bit Philox key the shim derives from R's generator after set.seed(r_seed) (two unif_rand() draws)."""
    return int(lib().rshim_seed_for(int(r_seed)))


def call(M, W, Z, P, E, Ap, Bp, Ae, Be, hyper, burn, eval, Pfixed=False, Zfixed=False, Thetafixed=False, Afixed=False,
         keep_par=True, r_seed=1):
    """.Call(`_signeR_GibbsSamplerCpp`, ...) with set.seed(r_seed).  Returns the reference's positional 8-slot list
    (None where it returns NA)."""
    M, W, Z, P, E, Ap, Bp, Ae, Be = (_f(x) for x in (M, W, Z, P, E, Ap, Bp, Ae, Be))
    I, J = M.shape
    w_cols = W.shape[1]
    K = P.shape[1]
    R = int(eval)
    f = dict(order="F")
    Zs = np.zeros((I, J, K, 1), **f); Ps = np.zeros((I, K, R), **f); Es = np.zeros((K, J, R), **f)
    Aps = np.zeros((I, K, R), **f); Bps = np.zeros((I, K, R), **f); Aes = np.zeros((K, J, R), **f); Bes = np.zeros((K, J, R), **f)
    hv = np.asarray(hyper, dtype=np.float64)
    err = C.create_string_buffer(1024)
    rc = lib().rshim_call(I, J, K, _p(M), _p(W), _p(Z), _p(P), _p(E), _p(Ap), _p(Bp), _p(Ae), _p(Be), _p(hv), int(burn), R,
                          int(Pfixed), int(Zfixed), int(Thetafixed), int(Afixed), int(keep_par), int(r_seed), int(w_cols),
                          _p(Zs), _p(Ps), _p(Es), _p(Aps), _p(Bps), _p(Aes), _p(Bes), err, 1024)
    if rc == 1:
        raise RError(err.value.decode("utf-8", "replace"))
    if rc != 0:
        raise AssertionError("the shim returned a malformed list (slots, NA pattern or dim names)")
    if keep_par:
        return [Zs, None, Ps, Es, Aps, Bps, Aes, Bes]
    return [None, None, Ps, Es, None, Bps, None, Bes]
