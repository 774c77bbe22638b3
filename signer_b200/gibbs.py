"""Host-side mirror of the reference's R<->native boundary for the Gibbs sampler.

`GibbsSamplerCpp` has the reference's name, its 24 positional arguments with the same meaning
(R/RcppExports.R:8-10; src/gibbs_2.cpp:253-263) and returns the same positional 8-element list
(src/gibbs_2.cpp:406-423; `None` where the reference returns NA), so callers written against the
reference (R/cpp_eBayesNMF.R:104-114,165-177) read the same slots.  R is not available in this
image, so this module plays the role of the Rcpp shim (r_shim/) in tests and benchmarks; both
call the same C ABI (include/signer_gibbs.h).

Arrays are float64 in Fortran (column-major) order, i.e. exactly R's / Armadillo's memory.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import Opts, Outputs, Problem, State, check, dp, ip

_RNG = np.random.default_rng()


def set_seed(seed):
    """Role of R's set.seed(): every later call draws its 64-bit Philox key from this generator,
    like the shim draws it from R's RNG inside an RNGScope (src/RcppExports.cpp:34)."""
    global _RNG
    _RNG = np.random.default_rng(seed)


def next_seed():
    return int(_RNG.integers(0, 2 ** 63 - 1, dtype=np.int64))


def _f64(a, shape=None, name="array"):
    a = np.asfortranarray(np.asarray(a, dtype=np.float64))
    if shape is not None and a.shape != tuple(shape):
        raise ValueError("%s has shape %s, expected %s" % (name, a.shape, tuple(shape)))
    return a


def _ptr(a):
    return None if a is None else a.ctypes.data_as(dp)


def release_caches():
    """signer_release_caches(): drop the cached cohorts, trim the device memory pools, free the cached pinned buffers."""
    _lib.lib().signer_release_caches()


class SampleList(list):
    """The reference's 8-slot return list plus what the device computed on the way
    (log-likelihoods, BICs, final state, final Z statistics)."""
    loglik = None
    bic = None
    state = None
    Zin = None
    Znj = None
    Phat = None          # device-side posterior medians of the normalised samples (when requested)
    Ehat = None
    seed = None


def _make_opts(hyper, burn, eval, Pfixed, Zfixed, Thetafixed, Afixed, keep_par, seed, sweep0, device, forced):
    o = Opts()
    (o.ap, o.bp, o.ae, o.be, o.lp, o.le, o.var_ap, o.var_ae) = [float(x) for x in hyper]
    o.burn, o.eval = int(burn), int(eval)
    o.Pfixed, o.Zfixed, o.Thetafixed, o.Afixed, o.keep_par = (int(bool(x)) for x in (Pfixed, Zfixed, Thetafixed, Afixed, keep_par))
    o.seed = int(seed)
    o.sweep0 = int(sweep0)
    o.device = int(device)
    o.forced_order = forced.ctypes.data_as(ip) if forced is not None else None
    return o


def _forced(order, nsweeps):
    if order is None:
        return None
    a = np.ascontiguousarray(np.asarray(order, dtype=np.int32).reshape(nsweeps, 7))
    return a


class _Buffers:
    """Caller-allocated outputs in the reference's shapes."""

    def __init__(self, I, J, K, R, keep_par, want_Z=True, samples=True, summaries=False):
        f = dict(order="F")
        self.Phat = np.zeros((I, K), **f) if (summaries and R > 0) else None
        self.Ehat = np.zeros((K, J), **f) if (summaries and R > 0) else None
        self.Zs = np.zeros((I, J, K, 1), **f) if (keep_par and want_Z and R > 0) else None
        self.Ps = np.zeros((I, K, R), **f) if samples else None
        self.Es = np.zeros((K, J, R), **f) if samples else None
        self.Aps = np.zeros((I, K, R), **f) if (keep_par and samples) else None
        self.Aes = np.zeros((K, J, R), **f) if (keep_par and samples) else None
        self.Bps = np.zeros((I, K, R), **f) if samples else None
        self.Bes = np.zeros((K, J, R), **f) if samples else None
        self.loglik = np.zeros(R)
        self.bic = np.zeros(R)
        self.P = np.zeros((I, K), **f); self.Ap = np.zeros((I, K), **f); self.Bp = np.zeros((I, K), **f)
        self.E = np.zeros((K, J), **f); self.Ae = np.zeros((K, J), **f); self.Be = np.zeros((K, J), **f)
        self.Zin = np.zeros((I, K), dtype=np.int32, order="F")
        self.Znj = np.zeros((K, J), dtype=np.int32, order="F")

    def struct(self):
        o = Outputs()
        for name in ("Zs", "Ps", "Es", "Aps", "Aes", "Bps", "Bes", "loglik", "bic", "P", "E", "Ap", "Bp", "Ae", "Be", "Phat", "Ehat"):
            setattr(o, name, _ptr(getattr(self, name)))
        o.Zin = self.Zin.ctypes.data_as(ip)
        o.Znj = self.Znj.ctypes.data_as(ip)
        return o

    def as_list(self, keep_par, seed):
        # src/gibbs_2.cpp:406-423
        if keep_par:
            res = SampleList([self.Zs, None, self.Ps, self.Es, self.Aps, self.Bps, self.Aes, self.Bes])
        else:
            res = SampleList([None, None, self.Ps, self.Es, None, self.Bps, None, self.Bes])
        res.loglik, res.bic = self.loglik, self.bic
        res.state = dict(P=self.P, E=self.E, Ap=self.Ap, Bp=self.Bp, Ae=self.Ae, Be=self.Be)
        res.Zin, res.Znj = self.Zin, self.Znj
        res.Phat, res.Ehat = self.Phat, self.Ehat
        res.seed = seed
        return res


def GibbsSamplerCpp(M, W, Z, P, E, Ap, Bp, Ae, Be, ap, bp, ae, be, lp, le, var_ap, var_ae, burn, eval,
                    Pfixed, Zfixed, Thetafixed, Afixed, keep_par, *, seed=None, sweep0=0, device=0,
                    forced_order=None, shard_devices=None):
    """Drop-in for the reference's GibbsSamplerCpp (src/gibbs_2.cpp:253-424).

    Inputs are not modified (the reference takes them by value, src/RcppExports.cpp:35-58).
    Z may be None: the first allocation is then drawn on the device.  Keyword-only extras have no
    counterpart in R: `seed` replaces the RNGScope (default: drawn from set_seed()'s generator).
    Unsupported flag sets raise SignerError(code 2) like the shim raises an R error.
    """
    M = _f64(M, name="M")
    if M.ndim != 2:
        raise ValueError("M must be a matrix")
    I, J = M.shape
    W = _f64(W, (I, J), "W")
    P = _f64(P, name="P")
    K = P.shape[1]
    P = _f64(P, (I, K), "P"); Ap = _f64(Ap, (I, K), "Ap"); Bp = _f64(Bp, (I, K), "Bp")
    E = _f64(E, (K, J), "E"); Ae = _f64(Ae, (K, J), "Ae"); Be = _f64(Be, (K, J), "Be")
    Zc = None if Z is None else _f64(Z, (I, J, K), "Z")
    if seed is None:
        seed = next_seed()
    burn, eval = int(burn), int(eval)
    forced = _forced(forced_order, burn + eval)
    pr = Problem(I, J, K, _ptr(M), _ptr(W))
    st = State(_ptr(Zc), _ptr(P), _ptr(E), _ptr(Ap), _ptr(Bp), _ptr(Ae), _ptr(Be))
    op = _make_opts((ap, bp, ae, be, lp, le, var_ap, var_ae), burn, eval, Pfixed, Zfixed, Thetafixed, Afixed,
                    keep_par, seed, sweep0, device, forced)
    buf = _Buffers(I, J, K, eval, bool(keep_par))
    out = buf.struct()
    if shard_devices is not None:           # one chain over several GPUs, genome columns sharded (include/signer_gibbs.h)
        dev = np.ascontiguousarray(np.asarray(shard_devices, dtype=np.int32))
        check(_lib.lib().signer_gibbs_run_sharded(C.byref(pr), C.byref(st), C.byref(op), C.byref(out), dev.ctypes.data_as(ip), dev.size))
    else:
        check(_lib.lib().signer_gibbs_run(C.byref(pr), C.byref(st), C.byref(op), C.byref(out)))
    return buf.as_list(bool(keep_par), seed)


class Chain:
    """Device-resident chain (signer_chain_* of the C ABI): M, W, the parameters and the Z
    statistics stay in HBM between calls; used by the EM loop, the benchmarks and the tests."""

    def __init__(self, M, W, P, E, Ap, Bp, Ae, Be, hyper, Z=None, Pfixed=False, seed=None, sweep0=0, device=0):
        M = _f64(M, name="M")
        I, J = M.shape
        W = _f64(W, (I, J), "W")
        P = _f64(P, name="P")
        K = P.shape[1]
        P = _f64(P, (I, K), "P"); Ap = _f64(Ap, (I, K), "Ap"); Bp = _f64(Bp, (I, K), "Bp")
        E = _f64(E, (K, J), "E"); Ae = _f64(Ae, (K, J), "Ae"); Be = _f64(Be, (K, J), "Be")
        Zc = None if Z is None else _f64(Z, (I, J, K), "Z")
        self.I, self.J, self.K = I, J, K
        self.seed = next_seed() if seed is None else int(seed)
        self.device = device
        pr = Problem(I, J, K, _ptr(M), _ptr(W))
        st = State(_ptr(Zc), _ptr(P), _ptr(E), _ptr(Ap), _ptr(Bp), _ptr(Ae), _ptr(Be))
        op = _make_opts(hyper, 0, 0, Pfixed, False, False, False, False, self.seed, sweep0, device, None)
        h = C.c_void_p()
        check(_lib.lib().signer_chain_create(C.byref(pr), C.byref(st), C.byref(op), C.byref(h)))
        self._h = h

    def set_stream(self, cuda_stream_ptr):
        check(_lib.lib().signer_chain_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)))

    def set_hyper(self, hyper):
        h = np.asarray(hyper, dtype=np.float64)
        check(_lib.lib().signer_chain_set_hyper(self._h, _ptr(h)))

    def run(self, burn=0, eval=0, keep_par=False, fetch=True, want_Z=True, samples=True, forced_order=None, summaries=False):
        """burn sweeps without storage then eval sweeps with storage.  fetch=False leaves everything
        on the device (asynchronous; call sync())."""
        forced = _forced(forced_order, burn + eval)
        fptr = forced.ctypes.data_as(ip) if forced is not None else None
        if not fetch:
            check(_lib.lib().signer_chain_run(self._h, burn, eval, int(keep_par), fptr, None))
            return None
        buf = _Buffers(self.I, self.J, self.K, eval, bool(keep_par), want_Z=want_Z, samples=samples, summaries=summaries)
        out = buf.struct()
        check(_lib.lib().signer_chain_run(self._h, burn, eval, int(keep_par), fptr, C.byref(out)))
        return buf.as_list(bool(keep_par), self.seed)

    def sync(self):
        check(_lib.lib().signer_chain_sync(self._h))

    def state(self):
        buf = _Buffers(self.I, self.J, self.K, 0, False, samples=False)
        out = buf.struct()
        check(_lib.lib().signer_chain_fetch_state(self._h, C.byref(out)))
        st = dict(P=buf.P, E=buf.E, Ap=buf.Ap, Bp=buf.Bp, Ae=buf.Ae, Be=buf.Be, Zin=buf.Zin, Znj=buf.Znj)
        return st

    def stats(self):
        s = np.zeros(8)
        check(_lib.lib().signer_chain_fetch_stats(self._h, _ptr(s)))
        return s

    def profile(self, enable=True):
        ms = np.zeros(4)
        n = np.zeros(4, dtype=np.int64)
        check(_lib.lib().signer_chain_profile(self._h, int(enable), _ptr(ms), n.ctypes.data_as(C.POINTER(C.c_int64))))
        names = ("z_alloc_fused", "p_family", "e_family", "loglik")
        return {k: dict(ms=float(ms[i]), launches=int(n[i])) for i, k in enumerate(names)}

    @property
    def sweeps(self):
        return int(_lib.lib().signer_chain_sweeps(self._h))

    def draw_counts(self, enable=True):
        """(categorical draws, conditional binomials) of step 1 since the last call; enable/disable counting."""
        cnt = (C.c_uint64 * 2)()
        check(_lib.lib().signer_chain_draw_counts(self._h, int(enable), cnt))
        return int(cnt[0]), int(cnt[1])

    @property
    def launches(self):
        return int(_lib.lib().signer_chain_launches(self._h))

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().signer_chain_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
