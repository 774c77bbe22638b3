"""Shared helpers of the parity tests: run the CUDA path (through the C ABI) and the oracle on the
same inputs and compare bit for bit."""
import numpy as np

from signer_b200 import synth


def small_problem(I=24, J=37, K=3, seed=1, burden=300.0, zero_frac=0.0):
    M, W, _, _ = synth.cohort(I, J, max(K, 1), burden=burden, seed=seed, opportunity=True)
    if zero_frac > 0:
        rng = np.random.default_rng(seed + 99)
        M = np.asfortranarray(np.where(rng.uniform(size=M.shape) < zero_frac, 0.0, M))
    st = synth.init_state(I, J, K, seed=seed + 1)
    return M, W, st


def bits(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    return a.view(np.uint64)


def assert_bits_equal(got, want, name):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (name, got.shape, want.shape)
    gb, wb = bits(np.asfortranarray(got).ravel(order="K")), bits(np.asfortranarray(want).ravel(order="K"))
    # NaNs must match as NaNs, everything else bit for bit
    both_nan = np.isnan(got.ravel(order="K")) & np.isnan(want.ravel(order="K"))
    bad = (gb != wb) & ~both_nan
    if bad.any():
        i = int(np.flatnonzero(bad)[0])
        raise AssertionError("%s: %d of %d entries differ; first at flat index %d: got %r want %r" % (
            name, int(bad.sum()), bad.size, i, got.ravel(order="K")[i], want.ravel(order="K")[i]))


def run_both(sgmod, oracle, M, W, st, K, hyper=None, burn=0, eval=1, Pfixed=False, keep_par=True, seed=7,
             sweep0=0, forced_order=None, Z=None, device=0):
    h = synth.hyper_tuple(hyper)
    got = sgmod.GibbsSamplerCpp(M, W, Z, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], *h, burn, eval,
                                Pfixed, False, False, False, keep_par, seed=seed, sweep0=sweep0, device=device,
                                forced_order=forced_order)
    hd = dict(zip(("ap", "bp", "ae", "be", "lp", "le", "var_ap", "var_ae"), h))
    want = oracle.gibbs_run(M, W, Z, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], hyper=hd, burn=burn,
                            eval=eval, Pfixed=Pfixed, keep_par=keep_par, seed=seed, sweep0=sweep0,
                            forced_order=forced_order)
    return got, want


def compare_all(got, want, keep_par=True, check_Z=True):
    """got: SampleList from the CUDA path; want: oracle dict."""
    assert_bits_equal(got[2], want["Ps"], "Ps")
    assert_bits_equal(got[3], want["Es"], "Es")
    assert_bits_equal(got[5], want["Bps"], "Bps")
    assert_bits_equal(got[7], want["Bes"], "Bes")
    if keep_par:
        assert_bits_equal(got[4], want["Aps"], "Aps")
        assert_bits_equal(got[6], want["Aes"], "Aes")
        if check_Z:
            assert_bits_equal(got[0][:, :, :, 0], want["Z"], "Zs")
    else:
        assert got[0] is None and got[4] is None and got[6] is None
    assert got[1] is None
    for k in ("P", "E", "Ap", "Bp", "Ae", "Be"):
        assert_bits_equal(got.state[k], want[k], "state " + k)
    assert np.array_equal(got.Zin, want["Zin"].astype(np.int64)), "Zin"
    assert np.array_equal(got.Znj, want["Znj"].astype(np.int64)), "Znj"
    assert_bits_equal(got.loglik, want["loglik"], "loglik")
    np.testing.assert_allclose(got.bic, want["bic"], rtol=1e-13, atol=0)
