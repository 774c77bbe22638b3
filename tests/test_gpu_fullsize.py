"""BASELINE.json's full-size workload (config 3: 96 x 10,000 genomes, K = 20) through size-independent properties --
the oracle would need ~1.3 s per sweep here, so the checks are invariants of the model and of the implementation:
checksums of the integer allocations against the data, non-negativity, run-to-run determinism (integer atomics and
dynamic work claiming must not leak into results), continuation, and finite/positive samples."""
import numpy as np
import pytest

from signer_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cfg3():
    M, W, _, _ = synth.named("cfg3_96x10000_K20")
    st = synth.init_state(96, 10000, 20, seed=20001)
    return M, W, st


def _chain(sg, cfg3, seed=77):
    M, W, st = cfg3
    return sg.Chain(M, W, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], synth.hyper_tuple(), seed=seed)


def test_allocation_checksums_at_full_size(sg, cfg3):
    M = np.trunc(cfg3[0]).astype(np.int64)
    with _chain(sg, cfg3) as c:
        for burn in (0, 1, 7):                                   # initial allocation, then after sweeps
            s = c.run(burn=burn, eval=0)
            Zin, Znj = s.Zin.astype(np.int64), s.Znj.astype(np.int64)
            assert Zin.min() >= 0 and Znj.min() >= 0
            assert np.array_equal(Zin.sum(axis=1), M.sum(axis=1))     # per context: sum_k sum_j Z = sum_j M
            assert np.array_equal(Znj.sum(axis=0), M.sum(axis=0))     # per genome : sum_k sum_i Z = sum_i M
            assert np.array_equal(Zin.sum(axis=0), Znj.sum(axis=1))   # per class  : both reductions of the same Z agree
        res = c.run(burn=0, eval=3, keep_par=True, want_Z=True)
    Z = res[0][:, :, :, 0]
    assert np.array_equal(Z.sum(axis=2), M) and Z.min() >= 0          # every cell: sum_k Z = M
    assert np.array_equal(Z.sum(axis=1), res.Zin) and np.array_equal(Z.sum(axis=0).T, res.Znj)
    for slot in (2, 3, 4, 5, 6, 7):
        assert np.all(np.isfinite(res[slot])) and np.all(res[slot] > 0)
    assert np.all(np.isfinite(res.bic)) and np.all(np.diff(res.loglik) != 0)


def test_determinism_and_continuation_at_full_size(sg, cfg3):
    with _chain(sg, cfg3) as c:
        a = c.run(burn=12, eval=0)
    with _chain(sg, cfg3) as c:
        b = c.run(burn=12, eval=0)
    with _chain(sg, cfg3) as c:
        c.run(burn=5, fetch=False)
        d = c.run(burn=7, eval=0)
    with _chain(sg, cfg3, seed=78) as c:
        e = c.run(burn=12, eval=0)
    for k in a.state:
        assert np.array_equal(a.state[k], b.state[k]), k          # same seed: bit-identical although atomics and scheduling vary
        assert np.array_equal(a.state[k], d.state[k]), k          # 5 + 7 sweeps == 12 sweeps
    assert np.array_equal(a.Zin, b.Zin) and np.array_equal(a.Znj, d.Znj)
    assert not np.array_equal(a.state["E"], e.state["E"])         # another seed: another chain


def test_wide_context_checksums(sg):
    """Config-4 rows (1536 contexts, K = 30) on a slice of genomes: the global-atomic Zin path."""
    M, W, _, _ = synth.cohort(1536, 600, 30, burden=20000.0, seed=20240504)
    st = synth.init_state(1536, 600, 30, seed=4)
    Mi = np.trunc(M).astype(np.int64)
    with sg.Chain(M, W, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], synth.hyper_tuple(), seed=5) as c:
        s = c.run(burn=3, eval=0)
    assert np.array_equal(s.Zin.astype(np.int64).sum(axis=1), Mi.sum(axis=1))
    assert np.array_equal(s.Znj.astype(np.int64).sum(axis=0), Mi.sum(axis=0))
