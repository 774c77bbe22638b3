"""C-ABI behaviour at the edges, through the Python host: degenerate shapes, zero-length runs, the EM statistics,
per-kernel profile and launch counters, hyper-parameter updates, streams."""
import numpy as np
import pytest

from helpers import assert_bits_equal, compare_all, run_both, small_problem
from signer_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("I,J,K", [(1, 1, 1), (1, 40, 2), (5, 1, 3), (2, 3, 2), (96, 31, 5), (33, 33, 33)])
def test_tiny_and_ragged_shapes(sg, oracle, I, J, K):
    M, W, st = small_problem(I=I, J=J, K=K, seed=I * 7 + J, burden=30.0 * I + 5)
    got, want = run_both(sg, oracle, M, W, st, K, burn=3, eval=2, seed=I + J + K)
    compare_all(got, want)


def test_zero_length_run_returns_the_input_state(sg):
    M, W, st = small_problem(I=10, J=12, K=3, seed=1)
    h = synth.hyper_tuple()
    r = sg.GibbsSamplerCpp(M, W, None, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], *h, 0, 0, False, False, False, False, True, seed=3)
    for k in ("P", "E", "Ap", "Bp", "Ae", "Be"):
        assert_bits_equal(r.state[k], st[k], k)
    assert r[2].shape == (10, 3, 0) and r.bic.shape == (0,)
    assert int(r.Zin.sum()) == int(np.trunc(M).sum())          # the initial allocation was drawn


def test_inputs_are_never_modified(sg):
    M, W, st = small_problem(I=12, J=20, K=3, seed=2)
    keep = {k: v.copy() for k, v in st.items()}
    M0, W0 = M.copy(), W.copy()
    sg.GibbsSamplerCpp(M, W, None, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], *synth.hyper_tuple(), 5, 3,
                       False, False, False, False, True, seed=1)
    assert np.array_equal(M, M0) and np.array_equal(W, W0)
    for k in st:
        assert np.array_equal(st[k], keep[k])


def test_em_statistics_equal_numpy_on_the_returned_cubes(sg):
    M, W, st = small_problem(I=20, J=40, K=4, seed=3)
    with sg.Chain(M, W, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], synth.hyper_tuple(), seed=4) as c:
        res = c.run(burn=10, eval=12, keep_par=True)
        s = c.stats()
    want = [res[4].sum(), res[5].sum(), np.log(res[5]).sum(), res[4].size, res[6].sum(), res[7].sum(), np.log(res[7]).sum(), res[6].size]
    np.testing.assert_allclose(s, want, rtol=1e-11)
    from signer_b200.estimate import EstimateParameters, estimate_from_stats
    np.testing.assert_allclose(estimate_from_stats(s[4:8], 1.0), EstimateParameters(res[6], res[7], 1.0, 1.0), rtol=1e-9)


def test_profile_and_launch_counters(sg):
    M, W, st = small_problem(I=20, J=40, K=4, seed=5)
    with sg.Chain(M, W, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], synth.hyper_tuple(), seed=6) as c:
        l0 = c.launches
        c.profile(True)
        c.run(burn=10, fetch=False)
        p = c.profile(False)
        assert c.launches - l0 == 30                          # three kernels per burn-in sweep
        assert p["z_alloc_fused"]["launches"] == 10 and p["p_family"]["launches"] == 10 and p["e_family"]["launches"] == 10
        assert all(p[k]["ms"] > 0 for k in ("z_alloc_fused", "p_family", "e_family"))
        c.run(eval=4, fetch=False); c.sync()
        assert c.sweeps == 14
        c.draw_counts(True)
        c.run(burn=2, fetch=False)
        cat, binom = c.draw_counts(False)
        assert cat + binom > 0 and cat <= 2 * int(np.trunc(M).sum())


def test_set_hyper_changes_the_draws_and_pfixed_guards(sg):
    M, W, st = small_problem(I=16, J=30, K=3, seed=7)
    h = synth.hyper_tuple()

    def final_bp(hyper):
        with sg.Chain(M, W, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], h, seed=8) as c:
            if hyper is not None:
                c.set_hyper(hyper)
            return c.run(burn=3, eval=0).state["Bp"]
    a, b, c2 = final_bp(None), final_bp(h), final_bp((5.0,) + h[1:])
    assert np.array_equal(a, b) and not np.array_equal(a, c2)
    with sg.Chain(M, W, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], h, seed=8, Pfixed=True) as c:
        s = c.run(burn=5, eval=0).state
    assert np.array_equal(s["P"], st["P"]) and np.array_equal(s["Bp"], st["Bp"]) and np.array_equal(s["Ap"], st["Ap"])


def test_chain_on_a_torch_stream(sg):
    torch = pytest.importorskip("torch")
    M, W, st = small_problem(I=16, J=30, K=3, seed=9)
    h = synth.hyper_tuple()
    stream = torch.cuda.Stream()
    with sg.Chain(M, W, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], h, seed=10) as c:
        c.set_stream(stream.cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); c.run(burn=20, fetch=False); e1.record(stream); stream.synchronize()
        assert e0.elapsed_time(e1) > 0
        c.set_stream(None)
        a = c.run(burn=0, eval=0).state
    with sg.Chain(M, W, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], h, seed=10) as c:
        b = c.run(burn=20, eval=0).state
    for k in a:
        assert np.array_equal(a[k], b[k])


def test_many_ring_chunks_are_bit_exact(sg, oracle, monkeypatch):
    """the samples leave the device ring in chunks through the copy-out pool while the chain keeps running: with a ring of
    two sweeps per set (15 and 7 chunks, both sets reused many times) every sample cube still equals the oracle's"""
    M, W, st = small_problem(I=40, J=300, K=6, seed=11)
    slice_bytes = 8 * (40 * 6 + 6 * 300)                            # one P-family + one E-family matrix
    monkeypatch.setenv("SIGNER_RING_BYTES", str(2 * 3 * slice_bytes))        # keep_par: three slices per sweep
    got, want = run_both(sg, oracle, M, W, st, 6, burn=4, eval=29, keep_par=True, seed=2211)
    compare_all(got, want)
    monkeypatch.setenv("SIGNER_RING_BYTES", str(2 * 2 * slice_bytes))        # keep_par=FALSE: two slices per sweep
    got, want = run_both(sg, oracle, M, W, st, 6, burn=2, eval=13, keep_par=False, seed=2212)
    compare_all(got, want, keep_par=False)


@pytest.mark.parametrize("e_minb,z_budget", [("2", "16"), ("3", "24"), ("2", "24"), ("3", "16")])
def test_register_budgets_give_identical_results(sg, oracle, monkeypatch, e_minb, z_budget):
    """z_alloc_fused and e_family exist with two register budgets each (chain_create picks by shape): the same code under a
    different register allocation, so every output must still equal the oracle bit for bit whichever build is forced"""
    M, W, st = small_problem(I=96, J=700, K=9, seed=12, burden=3000.0)
    M[3, 5] = 5000.0
    monkeypatch.setenv("SIGNER_E_MINB", e_minb)
    monkeypatch.setenv("SIGNER_Z_BUDGET", z_budget)
    got, want = run_both(sg, oracle, M, W, st, 9, burn=4, eval=3, keep_par=True, seed=3300 + int(e_minb) + int(z_budget))
    compare_all(got, want)
