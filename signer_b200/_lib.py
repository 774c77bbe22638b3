"""ctypes binding of libsigner_gibbs.so (include/signer_gibbs.h).

The library is the product; this module only loads it.  There is no fallback of any kind: if the
shared object is missing or cannot be loaded, importing a compute entry point raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsigner_gibbs.so")

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)

STATUS = {0: "ok", 1: "bad argument", 2: "unsupported flag combination", 3: "CUDA error",
          4: "out of device memory", 5: "NCCL error", 6: "non-finite input"}


class SignerError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("signer_gibbs: %s (%d): %s" % (STATUS.get(code, "?"), code, text))
        self.code = code


class Problem(C.Structure):
    _fields_ = [("I", C.c_int32), ("J", C.c_int32), ("K", C.c_int32), ("M", dp), ("W", dp)]


class State(C.Structure):
    _fields_ = [("Z", dp), ("P", dp), ("E", dp), ("Ap", dp), ("Bp", dp), ("Ae", dp), ("Be", dp)]


class Opts(C.Structure):
    _fields_ = [("ap", C.c_double), ("bp", C.c_double), ("ae", C.c_double), ("be", C.c_double),
                ("lp", C.c_double), ("le", C.c_double), ("var_ap", C.c_double), ("var_ae", C.c_double),
                ("burn", C.c_int32), ("eval", C.c_int32),
                ("Pfixed", C.c_int32), ("Zfixed", C.c_int32), ("Thetafixed", C.c_int32), ("Afixed", C.c_int32),
                ("keep_par", C.c_int32),
                ("seed", C.c_uint64), ("sweep0", C.c_uint32), ("device", C.c_int32),
                ("forced_order", ip)]


class Outputs(C.Structure):
    _fields_ = [("Zs", dp), ("Ps", dp), ("Es", dp), ("Aps", dp), ("Aes", dp), ("Bps", dp), ("Bes", dp),
                ("loglik", dp), ("bic", dp),
                ("P", dp), ("E", dp), ("Ap", dp), ("Bp", dp), ("Ae", dp), ("Be", dp),
                ("Zin", ip), ("Znj", ip), ("Phat", dp), ("Ehat", dp)]


class Task(C.Structure):
    _fields_ = [("K", C.c_int32), ("state", State), ("opts", Opts), ("loglik", dp), ("bic", dp),
                ("status", C.c_int32)]


EXPORTS = [
    "signer_gibbs_run", "signer_gibbs_run_sharded", "signer_chain_create", "signer_chain_set_stream", "signer_chain_set_hyper",
    "signer_chain_run", "signer_chain_sync", "signer_chain_fetch_state", "signer_chain_fetch_stats",
    "signer_chain_profile", "signer_chain_sweeps", "signer_chain_launches", "signer_chain_draw_counts", "signer_draw_ceiling", "signer_chain_destroy", "signer_gibbs_batch",
    "signer_test_math", "signer_test_binomial", "signer_test_gamma", "signer_test_perm",
    "signer_test_philox", "signer_release_caches", "signer_last_error", "signer_last_loop_ms", "signer_version", "signer_draw_spec_version", "signer_device_count",
]

_LIB = None


def lib():
    """Load libsigner_gibbs.so.  Raises if it has not been built (python __graft_entry__.py build)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with `make -C signer_b200/csrc` "
                          "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.signer_gibbs_run.restype = C.c_int
    L.signer_gibbs_run.argtypes = [C.POINTER(Problem), C.POINTER(State), C.POINTER(Opts), C.POINTER(Outputs)]
    L.signer_gibbs_run_sharded.restype = C.c_int
    L.signer_gibbs_run_sharded.argtypes = [C.POINTER(Problem), C.POINTER(State), C.POINTER(Opts), C.POINTER(Outputs), ip, C.c_int32]
    L.signer_chain_create.restype = C.c_int
    L.signer_chain_create.argtypes = [C.POINTER(Problem), C.POINTER(State), C.POINTER(Opts), C.POINTER(vp)]
    L.signer_chain_set_stream.restype = C.c_int
    L.signer_chain_set_stream.argtypes = [vp, vp]
    L.signer_chain_set_hyper.restype = C.c_int
    L.signer_chain_set_hyper.argtypes = [vp, dp]
    L.signer_chain_run.restype = C.c_int
    L.signer_chain_run.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, ip, C.POINTER(Outputs)]
    L.signer_chain_sync.restype = C.c_int
    L.signer_chain_sync.argtypes = [vp]
    L.signer_chain_fetch_state.restype = C.c_int
    L.signer_chain_fetch_state.argtypes = [vp, C.POINTER(Outputs)]
    L.signer_chain_fetch_stats.restype = C.c_int
    L.signer_chain_fetch_stats.argtypes = [vp, dp]
    L.signer_chain_profile.restype = C.c_int
    L.signer_chain_profile.argtypes = [vp, C.c_int32, dp, C.POINTER(C.c_int64)]
    L.signer_chain_sweeps.restype = C.c_int64
    L.signer_chain_sweeps.argtypes = [vp]
    L.signer_chain_launches.restype = C.c_int64
    L.signer_chain_launches.argtypes = [vp]
    L.signer_chain_draw_counts.restype = C.c_int
    L.signer_chain_draw_counts.argtypes = [vp, C.c_int32, C.POINTER(C.c_uint64)]
    L.signer_draw_ceiling.restype = C.c_int
    L.signer_draw_ceiling.argtypes = [C.c_int32, dp]
    L.signer_chain_destroy.restype = None
    L.signer_chain_destroy.argtypes = [vp]
    L.signer_gibbs_batch.restype = C.c_int
    L.signer_gibbs_batch.argtypes = [C.c_int32, C.c_int32, dp, dp, C.POINTER(Task), C.c_int32, ip, C.c_int32]
    L.signer_test_math.restype = C.c_int
    L.signer_test_math.argtypes = [C.c_int32, dp, dp, C.c_int32, C.c_int32]
    L.signer_test_binomial.restype = C.c_int
    L.signer_test_binomial.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_int32, ip, dp, ip, C.c_int32, C.c_int32]
    L.signer_test_gamma.restype = C.c_int
    L.signer_test_gamma.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, dp, dp, dp, C.c_int32, C.c_int32]
    L.signer_test_perm.restype = None
    L.signer_test_perm.argtypes = [C.c_uint64, C.c_uint32, ip]
    L.signer_test_philox.restype = None
    L.signer_test_philox.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.signer_release_caches.restype = None
    L.signer_release_caches.argtypes = []
    L.signer_last_error.restype = C.c_char_p
    L.signer_last_error.argtypes = []
    L.signer_last_loop_ms.restype = C.c_double
    L.signer_last_loop_ms.argtypes = []
    L.signer_version.restype = C.c_int
    L.signer_draw_spec_version.restype = C.c_int
    L.signer_device_count.restype = C.c_int
    _LIB = L
    return L


def check(rc):
    if rc != 0:
        raise SignerError(rc, lib().signer_last_error().decode("utf-8", "replace"))
