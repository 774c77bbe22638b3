"""Golden vectors of the draw specification (tests/golden/spec_v7.npz, made by make_golden.py):
the oracle must reproduce them here on the CPU, the CUDA path on the GPU."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import assert_bits_equal

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "spec_v7.npz"))
STATE = ("P", "E", "Ap", "Bp", "Ae", "Be")


def test_oracle_reproduces_golden_primitives(oracle):
    L = oracle.lib()
    n, p = G["binom_n"], G["binom_p"]
    got = np.array([L.oracle_binomial(99, 7, 1, e, 3, int(n[e]), float(p[e])) for e in range(n.size)])
    assert np.array_equal(got, G["binom_out"])
    got = np.array([L.oracle_gamma(98, 6, 5, e, float(G["gamma_shape"][e]), float(G["gamma_rate"][e])) for e in range(400)])
    assert_bits_equal(got, G["gamma_out"], "gamma")
    assert np.array_equal(np.array([oracle.perm(12345, t) for t in range(16)]), G["perms"])


def test_oracle_reproduces_golden_chain(oracle):
    st = {k: G["chain_init_" + k] for k in STATE}
    r = oracle.gibbs_run(G["chain_M"], G["chain_W"], None, *[st[k] for k in STATE], burn=6, eval=3, seed=4242)
    for k in ("Ps", "Es", "Aps", "Bps", "Aes", "Bes", "Z", "loglik", "Zin", "Znj"):
        assert_bits_equal(r[k], G["chain_" + k], k)


def test_oracle_reproduces_golden_wide_chain(oracle):
    """35 W E' segments: two levels of the blocked ordered sum (draw specification v7)"""
    st = {k: G["wide_init_" + k] for k in STATE}
    r = oracle.gibbs_run(G["wide_M"], G["wide_W"], None, *[st[k] for k in STATE], burn=2, eval=2, seed=4243)
    for k in ("Ps", "Es", "Bps", "Bes", "loglik", "Zin", "Znj"):
        assert_bits_equal(r[k], G["wide_" + k], k)


@pytest.mark.gpu
def test_cuda_reproduces_golden_wide_chain(sg):
    from signer_b200 import synth
    st = {k: G["wide_init_" + k] for k in STATE}
    got = sg.GibbsSamplerCpp(G["wide_M"], G["wide_W"], None, *[st[k] for k in STATE], *synth.hyper_tuple(), 2, 2,
                             False, False, False, False, True, seed=4243)
    assert_bits_equal(got[2], G["wide_Ps"], "Ps"); assert_bits_equal(got[3], G["wide_Es"], "Es")
    assert_bits_equal(got[5], G["wide_Bps"], "Bps"); assert_bits_equal(got[7], G["wide_Bes"], "Bes")
    assert_bits_equal(got.loglik, G["wide_loglik"], "loglik")
    assert np.array_equal(got.Zin, G["wide_Zin"].astype(np.int64)) and np.array_equal(got.Znj, G["wide_Znj"].astype(np.int64))


@pytest.mark.gpu
def test_cuda_reproduces_golden_primitives(sg):
    from signer_b200 import _lib
    L = _lib.lib()
    n = np.ascontiguousarray(G["binom_n"]); p = np.ascontiguousarray(G["binom_p"])
    out = np.zeros(n.size, dtype=np.int32)
    _lib.check(L.signer_test_binomial(99, 7, 1, 3, n.ctypes.data_as(_lib.ip), p.ctypes.data_as(_lib.dp),
                                      out.ctypes.data_as(_lib.ip), n.size, 0))
    assert np.array_equal(out, G["binom_out"])
    sh = np.ascontiguousarray(G["gamma_shape"]); ra = np.ascontiguousarray(G["gamma_rate"]); g = np.zeros(400)
    _lib.check(L.signer_test_gamma(98, 6, 5, sh.ctypes.data_as(_lib.dp), ra.ctypes.data_as(_lib.dp), g.ctypes.data_as(_lib.dp), 400, 0))
    assert_bits_equal(g, G["gamma_out"], "gamma")


@pytest.mark.gpu
def test_cuda_reproduces_golden_chain(sg):
    from signer_b200 import synth
    st = {k: G["chain_init_" + k] for k in STATE}
    h = synth.hyper_tuple()
    got = sg.GibbsSamplerCpp(G["chain_M"], G["chain_W"], None, *[st[k] for k in STATE], *h, 6, 3,
                             False, False, False, False, True, seed=4242)
    assert_bits_equal(got[2], G["chain_Ps"], "Ps"); assert_bits_equal(got[3], G["chain_Es"], "Es")
    assert_bits_equal(got[4], G["chain_Aps"], "Aps"); assert_bits_equal(got[5], G["chain_Bps"], "Bps")
    assert_bits_equal(got[6], G["chain_Aes"], "Aes"); assert_bits_equal(got[7], G["chain_Bes"], "Bes")
    assert_bits_equal(got[0][:, :, :, 0], G["chain_Z"], "Zs")
    assert_bits_equal(got.loglik, G["chain_loglik"], "loglik")
    assert np.array_equal(got.Zin, G["chain_Zin"].astype(np.int64)) and np.array_equal(got.Znj, G["chain_Znj"].astype(np.int64))
