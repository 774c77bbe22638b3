"""Bit-exact parity with the oracle AT THE SIZES BASELINE.json names -- not checksums: every output of a short chain.
These shapes stress what small ones cannot: the count-sorted work list at ~28 k batches with dynamic claiming, the
chain -> direct-remainder hand-off on heavy cells, BTRS on the pentanucleotide burdens, Znj atomics over L2 at full
width, the K = 40 shared-memory path, ring chunking of 3.2 MB slices.  The oracle needs ~1.3 s per config-3 sweep, so
the chains are short (an initial allocation plus 2-3 sweeps); total well under five minutes."""
import numpy as np
import pytest

from helpers import assert_bits_equal, compare_all, run_both
from signer_b200 import synth

pytestmark = pytest.mark.gpu


def test_config3_full_size_bit_exact(sg, oracle):
    """96 x 10,000, K = 20 (BASELINE.json configs[2]): initial allocation, 2 burn-in sweeps, 1 evaluation sweep --
    all six sample cubes, the final Z cube, the state, Zin/Znj, the log-likelihood."""
    M, W, _, _ = synth.named("cfg3_96x10000_K20")
    st = synth.init_state(96, 10000, 20, seed=20001)
    got, want = run_both(sg, oracle, M, W, st, 20, burn=2, eval=1, keep_par=True, seed=31)
    compare_all(got, want)
    assert np.array_equal(got[0][:, :, :, 0].sum(axis=2), np.trunc(M))


def test_config4_slice_bit_exact(sg, oracle):
    """1536 contexts x 1,000 genomes, K = 30, burden 20,000 (config 4's rows and counts): long conditional-binomial
    chains, BTRS, global-atomic Zin."""
    M, W, _, _ = synth.cohort(1536, 1000, 30, burden=20000.0, seed=20240504)
    st = synth.init_state(1536, 1000, 30, seed=4)
    got, want = run_both(sg, oracle, M, W, st, 30, burn=1, eval=1, keep_par=False, seed=32)
    compare_all(got, want, keep_par=False)


def test_config5_shape_bit_exact(sg, oracle):
    """96 x 20,000 genomes, K = 40 (config 5's widest candidate on a fifth of its genomes): one evaluation sweep."""
    M, W, _, _ = synth.cohort(96, 20000, 12, burden=2000.0, seed=20240505)
    st = synth.init_state(96, 20000, 40, seed=5)
    got, want = run_both(sg, oracle, M, W, st, 40, burn=0, eval=1, keep_par=False, seed=33)
    compare_all(got, want, keep_par=False)


@pytest.mark.parametrize("I,J,K", [(96, 97, 95), (130, 140, 128), (96, 130, 88)])
def test_largest_k_with_evaluation_sweeps(sg, oracle, I, J, K):
    """K up to 128 is accepted; the default signeR() search always evaluates Nmax = min(dim(M)) - 1 = 95 on 96 contexts
    (R/signeR.R:150, R/Optimal_sigs.R:15).  The log-likelihood kernel needs more than 48 KB of shared memory from K = 88."""
    M, W, _, _ = synth.cohort(I, J, 6, burden=40.0 * I, seed=K)
    st = synth.init_state(I, J, K, seed=K + 1)
    got, want = run_both(sg, oracle, M, W, st, K, burn=1, eval=2, keep_par=True, seed=34)
    compare_all(got, want)


def test_no_burn_in_large_outputs_are_intact(sg, oracle):
    """burn = 0 and outputs of more than 8 MB each: the copy-out starts while the helper threads are still touching the
    caller's pages -- every byte of every cube must still equal the oracle's (a page touch must never undo a copy)."""
    M, W, _, _ = synth.cohort(96, 400, 5, burden=300.0, seed=77)
    K = 8
    st = synth.init_state(96, 400, K, seed=78)
    ev = 330                                               # Es, Bes, Aes: 330 x 25.6 KB = 8.4 MB each
    assert 8 * K * 400 * ev > (8 << 20)
    got, want = run_both(sg, oracle, M, W, st, K, burn=0, eval=ev, keep_par=True, seed=35)
    compare_all(got, want)


def test_em_statistics_do_not_depend_on_the_ring_size(sg, monkeypatch):
    """eBayesNMF(estimate_hyper=TRUE) runs EM_eval_it sweeps with the samples kept on the device and asks for their
    sufficient statistics (R/EstimateParameters.R:1-13).  With a ring of two sweeps per set the 40 evaluation sweeps wrap
    the rings ten times; the statistics must still cover all of them."""
    M, W, _, _ = synth.cohort(24, 60, 3, burden=600.0, seed=5)
    K = 3
    st = synth.init_state(24, 60, K, seed=6)
    h = synth.hyper_tuple()
    slice_bytes = 8 * (24 * K + K * 60)
    with sg.Chain(M, W, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], h, seed=9) as c:
        res = c.run(burn=5, eval=40, keep_par=True)
    want = [res[4].sum(), res[5].sum(), np.log(res[5]).sum(), res[4].size, res[6].sum(), res[7].sum(), np.log(res[7]).sum(), res[6].size]
    monkeypatch.setenv("SIGNER_RING_BYTES", str(2 * 3 * slice_bytes))
    with sg.Chain(M, W, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], h, seed=9) as c:
        c.run(burn=5, eval=40, keep_par=True, fetch=False)              # everything stays on the device (the EM loop's call)
        s = c.stats()
        c.run(burn=1, eval=3, keep_par=False, fetch=False)              # a later run without keep_par invalidates them
        with pytest.raises(Exception):
            c.stats()
    np.testing.assert_allclose(s, want, rtol=1e-11)
    # the host mirror of the EM loop end to end under the same small ring
    r = sg.eBayesNMF(M, W, K, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 10.0, 10.0, start="random", burn_it=20, eval_it=20, estimate_hyper=True,
                     EM_lim=3, EM_eval_it=30, rng=np.random.default_rng(1))
    assert np.all(np.isfinite(r[2]))
