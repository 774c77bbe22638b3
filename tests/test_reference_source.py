"""The reference's OWN source (src/gibbs_2.cpp, compiled unmodified by oracle/ref_build against stand-in
Rcpp/Armadillo/R headers whose RNG and special functions are the draw specification) against the oracle's
restatement: every sample cube and the final Z must agree bit for bit.  This pins the restatement of the
update equations, clamps, loop orders, guards, storage rules and return list to the real source; it
cannot pin R's nmath, which is absent.  Skipped where oracle/_ref was not built (needs /root/reference at
build time; the built library travels to the GPU box)."""
import numpy as np
import pytest

from helpers import assert_bits_equal
from signer_b200 import synth

ref_runner = pytest.importorskip("oracle.ref_runner")
pytestmark = pytest.mark.skipif(not ref_runner.available(), reason="oracle/_ref not built")


def _run_both(oracle, M, W, st, burn, ev, Pfixed=False, keep_par=True, seed=3):
    Z0 = oracle.gibbs_run(M, W, None, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], burn=0, eval=0, seed=seed)["Z"]
    a = oracle.gibbs_run(M, W, Z0, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], burn=burn, eval=ev, Pfixed=Pfixed,
                         keep_par=keep_par, seed=seed)
    b = ref_runner.gibbs_run(M, W, Z0, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], burn=burn, eval=ev, Pfixed=Pfixed,
                             keep_par=keep_par, seed=seed)
    return a, b


@pytest.mark.parametrize("I,J,K", [(12, 21, 4), (96, 21, 5), (30, 140, 3), (8, 5, 1)])
def test_reference_source_equals_oracle(oracle, I, J, K):
    M, W, _, _ = synth.cohort(I, J, max(K, 1), burden=60.0 * I, seed=I + J)
    M[0, 0] = 3000.0
    st = synth.init_state(I, J, K, seed=7)
    a, b = _run_both(oracle, M, W, st, burn=12, ev=5)
    assert b["list_shape_ok"]                                   # 8 slots, NA where expected, named dims i,n,r / n,j,r
    for k in ("Ps", "Es", "Aps", "Bps", "Aes", "Bes", "Z"):
        assert_bits_equal(b[k], a[k], k)
    assert b["max_prob_dev"] < 1e-14                            # Fi of the reference == P o E / S of the specification


def test_reference_source_pfixed_and_keep_par_false(oracle):
    M, W, cosmic = synth.breast21()
    st = synth.init_state(96, 21, 6, seed=3, fixedP=cosmic)
    a, b = _run_both(oracle, M, W, st, burn=6, ev=4, Pfixed=True, keep_par=False)
    assert b["list_shape_ok"]
    for k in ("Ps", "Es", "Bps", "Bes"):
        assert_bits_equal(b[k], a[k], k)
    assert not b["Aps"].any() and not b["Aes"].any()            # NA slots were not filled
