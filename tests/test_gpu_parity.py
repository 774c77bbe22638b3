"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs.  Bar: bit-exact for the integer allocations AND for every fp64 output, because the
draw specification only uses correctly rounded operations in a fixed order."""
import ctypes as C

import numpy as np
from scipy import stats
import pytest

from helpers import assert_bits_equal, compare_all, run_both, small_problem
from signer_b200 import _lib, synth

pytestmark = pytest.mark.gpu


def _dp(a):
    return a.ctypes.data_as(_lib.dp)


def test_device_present(sg):
    assert _lib.lib().signer_device_count() >= 1


@pytest.mark.parametrize("fn,name", [(0, "oracle_log"), (1, "oracle_exp"), (2, "oracle_lgamma"), (3, "oracle_normal")])
def test_device_math_bit_exact(sg, oracle, fn, name):
    rng = np.random.default_rng(fn)
    if fn == 0:
        x = np.concatenate([10 ** rng.uniform(-320, 308, 20000), rng.uniform(0.4, 2.5, 20000), [0.0, 1.0, 5e-324, np.inf, -1.0, np.nan]])
    elif fn == 1:
        x = np.concatenate([rng.uniform(-760, 720, 30000), rng.uniform(-1, 1, 10000), [0.0, -745.2, 709.8, 1e-300, np.nan, -np.inf]])
    elif fn == 2:
        x = np.concatenate([10 ** rng.uniform(-320, 300, 20000), rng.uniform(0, 30, 20000), [0.0, 1.0, 2.0, 8.0, 5e-324, np.inf]])
    else:
        x = rng.integers(0, 4, 120000).astype(np.float64)        # attempt numbers; 120000 streams reach wedges and the tail
    x = np.ascontiguousarray(x)
    y = np.zeros_like(x)
    _lib.check(_lib.lib().signer_test_math(fn, _dp(x), _dp(y), x.size, 0))
    f = getattr(oracle.lib(), name)
    if fn == 3:
        want = np.array([f(12345, 0, 3, e, int(v)) for e, v in enumerate(x)])
        assert np.abs(want).max() > 3.4426 and stats.kstest(want, "norm").pvalue > 1e-4   # the tail beyond R was exercised
    else:
        want = np.array([f(float(v)) for v in x])
    assert_bits_equal(y, want, name)


def test_binomial_bit_exact(sg, oracle):
    rng = np.random.default_rng(5)
    n = np.concatenate([rng.integers(1, 60, 20000), rng.integers(1, 30000, 20000), [1, 1, 2, 2147483647 // 4]]).astype(np.int32)
    p = np.concatenate([10 ** rng.uniform(-9, 0, 20000), rng.uniform(0, 1, 20000), [0.5, 1e-300, 0.999999, 1e-7]])
    p = np.clip(p, 1e-300, 1 - 1e-16)
    L = oracle.lib()
    for kstep in (0, 1, 6, 127):
        out = np.zeros(n.size, dtype=np.int32)
        _lib.check(_lib.lib().signer_test_binomial(11, 3, 1, kstep, n.ctypes.data_as(_lib.ip), _dp(p), out.ctypes.data_as(_lib.ip), n.size, 0))
        want = np.array([L.oracle_binomial(11, 3, 1, e, kstep, int(n[e]), float(p[e])) for e in range(n.size)])
        assert np.array_equal(out, want)
    assert (out >= 0).all() and (out <= n).all()


def test_gamma_bit_exact(sg, oracle):
    rng = np.random.default_rng(6)
    shape = np.concatenate([10 ** rng.uniform(-170, 4, 20000), rng.uniform(0.5, 50, 20000), [1.0, 1e-160, 0.999999, 1e5]])
    rate = np.concatenate([10 ** rng.uniform(-10, 10, 40000), [1.0, 1e170, 1e-170, 3.0]])
    out = np.zeros_like(shape)
    _lib.check(_lib.lib().signer_test_gamma(13, 2, 5, _dp(shape), _dp(rate), _dp(out), shape.size, 0))
    L = oracle.lib()
    want = np.array([L.oracle_gamma(13, 2, 5, e, float(shape[e]), float(rate[e])) for e in range(shape.size)])
    assert_bits_equal(out, want, "gamma")


def test_perm_and_philox_match_oracle(sg, oracle):
    L = _lib.lib()
    for seed in (0, 1, 2 ** 63 + 12345):
        for sweep in (0, 1, 77, 2 ** 32 - 1):
            o = (C.c_int32 * 7)()
            L.signer_test_perm(seed, sweep, o)
            assert list(o) == oracle.perm(seed, sweep)


@pytest.mark.parametrize("step", [1, 2, 3, 4, 5, 6, 7])
def test_teacher_forced_single_step(sg, oracle, step):
    M, W, st = small_problem(I=40, J=150, K=6, seed=step)
    got, want = run_both(sg, oracle, M, W, st, 6, burn=0, eval=1, forced_order=[step, 0, 0, 0, 0, 0, 0], seed=100 + step)
    compare_all(got, want, check_Z=(step == 1))


@pytest.mark.parametrize("I,J,K", [(96, 21, 5), (33, 129, 7), (7, 5, 1), (64, 257, 4), (100, 300, 13)])
def test_free_running_chain_bit_exact(sg, oracle, I, J, K):
    M, W, st = small_problem(I=I, J=J, K=K, seed=I + J + K, burden=40.0 * I)
    got, want = run_both(sg, oracle, M, W, st, K, burn=25, eval=6, keep_par=True, seed=I * 1000 + J)
    compare_all(got, want)
    tot = M.astype(np.int64).sum()
    assert got.Zin.sum() == tot and got.Znj.sum() == tot
    assert np.array_equal(got[0][:, :, :, 0].sum(axis=2), np.trunc(M))
    if K == 1:
        assert np.array_equal(got[0][:, :, 0, 0], np.trunc(M))


def test_breast21_fixture_chain(sg, oracle):
    M, W, _ = synth.breast21()
    st = synth.init_state(96, 21, 5, seed=42)
    got, want = run_both(sg, oracle, M, W, st, 5, burn=40, eval=10, seed=42)
    compare_all(got, want)


def test_keep_par_false_and_pfixed(sg, oracle):
    M, W, cosmic = synth.breast21()
    st = synth.init_state(96, 21, 6, seed=3, fixedP=cosmic)
    got, want = run_both(sg, oracle, M, W, st, 6, burn=10, eval=5, Pfixed=True, keep_par=False, seed=9)
    compare_all(got, want, keep_par=False)
    # with P fixed every Ps slice is P and Bps stays what was passed in (zeros), src/gibbs_2.cpp:311-315,325-333
    for r in range(5):
        assert np.array_equal(got[2][:, :, r], cosmic)
        assert not got[5][:, :, r].any()


def test_given_Z_cube_is_reduced_on_entry(sg, oracle):
    M, W, st = small_problem(I=20, J=45, K=4, seed=8)
    # a valid initial allocation: everything in class 0
    Z = np.zeros((20, 45, 4), order="F")
    Z[:, :, 0] = np.trunc(M)
    # step 2 first: it must see sum_j Z of the GIVEN cube
    got, want = run_both(sg, oracle, M, W, st, 4, burn=0, eval=2, Z=Z, forced_order=[2, 3, 1, 4, 5, 6, 7] * 2, seed=21)
    compare_all(got, want)


def test_zero_counts_and_heavy_cells(sg, oracle):
    M, W, st = small_problem(I=32, J=64, K=5, seed=12, burden=50.0, zero_frac=0.6)
    M[3, 7] = 18171.0   # heaviest cell of the vignette data; forces the BTRS branch
    M[0, 0] = 250000.0
    got, want = run_both(sg, oracle, M, W, st, 5, burn=5, eval=3, seed=33)
    compare_all(got, want)
    M0 = np.zeros_like(M)
    got, want = run_both(sg, oracle, M0, W, st, 5, burn=2, eval=2, seed=34)
    compare_all(got, want)
    assert got.Zin.sum() == 0


def test_continuation_equals_one_run(sg, oracle):
    """Chain handle: 10 + 10 sweeps give the same state as 20 (sweep counters continue)."""
    M, W, st = small_problem(I=30, J=70, K=4, seed=2)
    h = synth.hyper_tuple()
    with sg.Chain(M, W, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], h, seed=5) as c:
        c.run(burn=10, fetch=False)
        a = c.run(burn=10, eval=0)
        assert c.sweeps == 20
    with sg.Chain(M, W, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], h, seed=5) as c:
        b = c.run(burn=20, eval=0)
    for k in a.state:
        assert_bits_equal(a.state[k], b.state[k], k)
    assert np.array_equal(a.Zin, b.Zin) and np.array_equal(a.Znj, b.Znj)


def test_errors(sg):
    M, W, st = small_problem(I=8, J=9, K=2, seed=1)
    h = synth.hyper_tuple()
    args = (st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], *h, 1, 1)
    with pytest.raises(sg.SignerError) as e:
        sg.GibbsSamplerCpp(M, W, None, *args, False, True, False, False, True)
    assert e.value.code == 2
    with pytest.raises(sg.SignerError) as e:
        sg.GibbsSamplerCpp(M, W, None, *args, False, False, False, True, True)
    assert e.value.code == 2
    Mbad = M.copy(); Mbad[0, 0] = np.nan
    with pytest.raises(sg.SignerError) as e:
        sg.GibbsSamplerCpp(Mbad, W, None, *args, False, False, False, False, True)
    assert e.value.code == 6
    Wbad = W.copy(); Wbad[1, 1] = -1.0
    with pytest.raises(sg.SignerError) as e:
        sg.GibbsSamplerCpp(M, Wbad, None, *args, False, False, False, False, True)
    assert e.value.code == 6
    with pytest.raises(ValueError):
        sg.GibbsSamplerCpp(M, W[:, :-1], None, *args, False, False, False, False, True)


def test_cuda_equals_reference_source(sg, oracle):
    """The CUDA path against the reference's own src/gibbs_2.cpp (oracle/_ref, built from /root/reference in the
    dev container with the draw specification standing in for R nmath / Armadillo RNG)."""
    ref_runner = pytest.importorskip("oracle.ref_runner")
    if not ref_runner.available():
        pytest.skip("oracle/_ref not built")
    M, W, _ = synth.breast21()
    st = synth.init_state(96, 21, 5, seed=31)
    h = synth.hyper_tuple()
    Z0 = oracle.gibbs_run(M, W, None, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], burn=0, eval=0, seed=17)["Z"]
    got = sg.GibbsSamplerCpp(M, W, Z0, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], *h, 15, 6,
                             False, False, False, False, True, seed=17)
    ref = ref_runner.gibbs_run(M, W, Z0, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], burn=15, eval=6, seed=17)
    assert ref["list_shape_ok"]
    for slot, k in ((2, "Ps"), (3, "Es"), (4, "Aps"), (5, "Bps"), (6, "Aes"), (7, "Bes")):
        assert_bits_equal(got[slot], ref[k], k)
    assert_bits_equal(got[0][:, :, :, 0], ref["Z"], "Zs")


@pytest.mark.parametrize("I,J,K,note", [
    (40, 70, 33, "K odd and > 32: scalar E loads, no touched-class mask"),
    (300, 60, 40, "I*K*4 = 48 KB > 40 KB: Zin accumulated with global atomics, not in shared memory"),
    (1536, 40, 30, "pentanucleotide contexts (config 4 rows)"),
    (96, 300, 64, "K = 64"),
])
def test_large_k_and_wide_context_paths(sg, oracle, I, J, K, note):
    M, W, st = small_problem(I=I, J=J, K=K, seed=K, burden=20.0 * I)
    M[0, 0] = 900.0; M[1, 1] = 40000.0
    got, want = run_both(sg, oracle, M, W, st, K, burn=4, eval=2, keep_par=True, seed=K * 7)
    compare_all(got, want)
