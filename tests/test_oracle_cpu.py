"""CPU tests of the oracle (the checker itself): third-party algorithms against their published
known answers, the deterministic math against libm/scipy, the samplers against exact laws, and the
invariants of the model.  'Parity unpinned' (no upstream vectors exist): these are what pins it."""
import math
import os

import numpy as np
import pytest
from scipy import special, stats

from signer_b200 import synth


def test_philox_known_answers(oracle):
    # Random123 kat_vectors, philox4x32-10
    assert oracle.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_uniform_maps(oracle):
    L = oracle.lib()
    assert L.oracle_u53(0, 0) == 2.0 ** -53 and L.oracle_u53(0xffffffff, 0xffffffff) == 1.0 - 2.0 ** -53
    L.oracle_u32.restype = __import__("ctypes").c_double
    L.oracle_u32.argtypes = [__import__("ctypes").c_uint32]
    assert L.oracle_u32(0) == 2.0 ** -33 and L.oracle_u32(0xffffffff) == 1.0 - 2.0 ** -33


def test_deterministic_math_accuracy(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(1)
    x = np.concatenate([10 ** rng.uniform(-320, 308, 4000), rng.uniform(0.5, 2, 4000), [1.0, 5e-324]])
    got = np.array([L.oracle_log(v) for v in x]); ref = np.log(x)
    assert np.max(np.abs(got - ref) / np.spacing(np.abs(ref) + 1e-300)) <= 2.0
    x = np.concatenate([rng.uniform(-700, 700, 6000), rng.uniform(-1, 1, 2000)])
    got = np.array([L.oracle_exp(v) for v in x]); ref = np.exp(x)
    assert np.max(np.abs(got - ref) / np.spacing(ref)) <= 2.0
    x = np.concatenate([10 ** rng.uniform(-300, 300, 3000), rng.uniform(0, 30, 3000)])
    got = np.array([L.oracle_lgamma(v) for v in x]); ref = special.gammaln(x)
    assert np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref))) < 2e-14
    assert L.oracle_log(0.0) == -math.inf and math.isnan(L.oracle_log(-1.0)) and L.oracle_exp(-800.0) == 0.0
    assert L.oracle_exp(800.0) == math.inf and L.oracle_lgamma(0.0) == math.inf


def test_normal_law(oracle):
    """ziggurat normal: KS against the normal cdf, moments, and the mass beyond a few thresholds (layers, wedges, tail)"""
    L = oracle.lib()
    z = np.array([L.oracle_normal(7, 1, 3, e, e & 3) for e in range(200000)])
    assert stats.kstest(z, "norm").pvalue > 1e-3
    assert abs(z.mean()) < 4 / np.sqrt(z.size) and abs(z.var() - 1) < 4 * np.sqrt(2 / z.size)
    for thr in (0.5, 1.0, 2.0, 3.0, 3.442619855899):
        pr = 2 * stats.norm.sf(thr)
        assert abs((np.abs(z) > thr).mean() - pr) < 4.5 * np.sqrt(pr * (1 - pr) / z.size), thr
    assert abs((z > 0).mean() - 0.5) < 4 / np.sqrt(z.size)


@pytest.mark.parametrize("n,p", [(1, 0.3), (20, 0.1), (50, 0.59), (120, 0.2), (200, 0.4), (1000, 0.02), (1000, 0.031),
                                 (18171, 0.37), (30, 1e-5), (1000, 0.9999)])
def test_binomial_law(oracle, n, p):
    """chi-square of the oracle's binomial (inversion below n*min(p,1-p) < 30, BTRS above) against the exact pmf."""
    L = oracle.lib()
    N = 12000
    x = np.array([L.oracle_binomial(7, 3, 1, e, e % 19, n, p) for e in range(N)])
    assert x.min() >= 0 and x.max() <= n
    lo = int(max(0, stats.binom.ppf(2e-4, n, p))); hi = int(min(n, stats.binom.ppf(1 - 2e-4, n, p)))
    if hi <= lo:
        assert abs(x.mean() - n * p) < 6 * math.sqrt(n * p * (1 - p) / N) + 1e-12
        return
    obs = np.array([(x <= lo).sum()] + [(x == k).sum() for k in range(lo + 1, hi)] + [(x >= hi).sum()])
    pr = np.array([stats.binom.cdf(lo, n, p)] + [stats.binom.pmf(k, n, p) for k in range(lo + 1, hi)] + [stats.binom.sf(hi - 1, n, p)])
    m = pr * N > 5
    chi = ((obs[m] - pr[m] * N) ** 2 / (pr[m] * N)).sum() + (obs[~m].sum() - pr[~m].sum() * N) ** 2 / max(pr[~m].sum() * N, 1e-9) * (pr[~m].sum() * N > 1)
    assert stats.chi2.sf(chi, max(int(m.sum()), 2) - 1) > 1e-4


@pytest.mark.parametrize("a,r", [(0.05, 2.0), (0.5, 1.0), (1.0, 3.0), (2.5, 0.5), (30.0, 1e4), (1e4, 1.0)])
def test_gamma_law(oracle, a, r):
    L = oracle.lib()
    g = np.array([L.oracle_gamma(11, 5, 3, e, a, r) for e in range(12000)])
    assert stats.kstest(g, lambda v: stats.gamma.cdf(v, a, scale=1 / r)).pvalue > 1e-4


def test_gamma_clamps_like_reference(oracle):
    L = oracle.lib()
    # shape is clamped at 1e-160 (src/gibbs_2.cpp:40): the draw underflows to 0, callers clamp it back to 1e-160
    assert L.oracle_gamma(1, 0, 6, 0, 0.0, 1.0) == 0.0
    # scale = max(1e-160, 1/rate) (:41): an infinite rate does not produce 0 or NaN
    g = L.oracle_gamma(1, 0, 6, 0, 2.0, math.inf)
    assert 0 < g < 1e-150


def test_multinomial_cell_laws_and_invariants(oracle):
    P = np.array([[0.1, 0.3, 0.0, 0.5, 0.1]]); E = np.array([[2.0], [1.0], [5.0], [0.5], [3.0]])
    q = P[0] * E[:, 0]; q = q / q.sum()
    for n in (7, 32, 33, 400):                       # direct method (<= 32) and chain (> 32)
        N = 4000
        acc = np.zeros(5)
        for sw in range(N):
            z = oracle.multinomial_cell(3, sw, 1, 0, 0, n, P, E)
            assert z.sum() == n and z[2] == 0 and (z >= 0).all()
            acc += z
        chi = ((acc - N * n * q)[q > 0] ** 2 / (N * n * q[q > 0])).sum()
        assert stats.chi2.sf(chi, 3) > 1e-4, (n, acc / N / n, q)
    assert oracle.multinomial_cell(3, 0, 1, 0, 0, 0, P, E).sum() == 0
    assert oracle.multinomial_cell(3, 0, 1, 0, 0, 17.9, P, E).sum() == 17          # truncation like `int size`
    # K = 1: Z = M ; degenerate probabilities: everything to the last class
    assert oracle.multinomial_cell(3, 0, 1, 0, 0, 9, np.array([[2.0]]), np.array([[3.0]]))[0] == 9
    z = oracle.multinomial_cell(3, 0, 1, 0, 0, 9, np.zeros((1, 3)), np.ones((3, 1)))
    assert list(z) == [0, 0, 9]


def test_multinomial_chain_joint_law(oracle):
    """large counts (bucket-ordered chain + direct draws for the remainder): every class marginal and every pair sum
    is binomial -- which pins the joint law well beyond the means"""
    rng = np.random.default_rng(3)
    K = 9
    P = rng.gamma(0.3, 1.0, size=(1, K)); P[0, 4] = 0.0
    E = rng.gamma(1.0, 1.0, size=(K, 1))
    q = P[0] * E[:, 0]; q = q / q.sum()
    for n in (40, 150, 3000):
        N = 3000
        Z = np.array([oracle.multinomial_cell(11, sw, 1, 0, 0, n, P, E) for sw in range(N)])
        assert (Z.sum(axis=1) == n).all() and (Z[:, 4] == 0).all()
        for a in range(K):
            for b in range(a, K):
                pr = q[a] if a == b else q[a] + q[b]
                if pr <= 0 or pr >= 1:
                    continue
                x = Z[:, a] if a == b else Z[:, a] + Z[:, b]
                zs = (x.mean() - n * pr) / np.sqrt(n * pr * (1 - pr) / N)
                assert abs(zs) < 4.5, (n, a, b, zs)
                vr = x.var() / (n * pr * (1 - pr))
                assert abs(vr - 1) < 5 * np.sqrt((2 + 1 / (n * pr * (1 - pr))) / N) + 0.01, (n, a, b, vr)   # var of s^2 incl. kurtosis


def test_perm_is_uniform(oracle):
    cnt = np.zeros((7, 7))
    for t in range(7000):
        o = oracle.perm(99, t)
        assert sorted(o) == [1, 2, 3, 4, 5, 6, 7]
        for pos, s in enumerate(o):
            cnt[pos, s - 1] += 1
    assert stats.chi2.sf(((cnt - 1000) ** 2 / 1000).sum(), 36) > 1e-4


def test_chain_invariants_and_loglik(oracle):
    M, W, _, _ = synth.cohort(20, 33, 3, burden=400.0, seed=3)
    st = synth.init_state(20, 33, 3, seed=4)
    r = oracle.gibbs_run(M, W, None, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], burn=8, eval=4, seed=5)
    Mi = np.trunc(M)
    assert np.array_equal(r["Z"].sum(axis=2), Mi)                  # sum_k Z = M in every cell
    assert np.array_equal(r["Zin"], r["Z"].sum(axis=1)) and np.array_equal(r["Znj"], r["Z"].sum(axis=0).T)
    assert r["Zin"].sum() == Mi.sum()
    for k in ("Ps", "Es", "Bps", "Bes", "Aps", "Aes"):
        assert np.all(np.isfinite(r[k])) and np.all(r[k] > 0)
    lam = (r["Ps"][:, :, -1] @ r["Es"][:, :, -1]) * W
    ll = np.sum(Mi * np.log(lam) - special.gammaln(Mi + 1) - lam)       # R/cpp_eBayesNMF.R:181-185
    assert abs(r["loglik"][-1] - ll) <= 1e-10 * abs(ll)
    assert np.allclose(r["bic"], 2 * r["loglik"] - 3 * (20 + 33) * math.log(33), rtol=1e-14)
    # the last sample slices are the final state (what R re-reads, R/cpp_eBayesNMF.R:115-139)
    assert np.array_equal(r["Ps"][:, :, -1], r["P"]) and np.array_equal(r["Es"][:, :, -1], r["E"])


def test_pfixed_and_forced_order(oracle):
    M, W, cosmic = synth.breast21()
    st = synth.init_state(96, 21, 6, seed=3, fixedP=cosmic)
    r = oracle.gibbs_run(M, W, None, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], burn=3, eval=2, Pfixed=True, seed=1)
    assert np.array_equal(r["P"], cosmic) and not r["Bp"].any() and not r["Ap"].any()
    # an empty forced order leaves the state untouched
    r2 = oracle.gibbs_run(M, W, None, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], burn=0, eval=1, seed=1,
                          forced_order=[0] * 7)
    assert np.array_equal(r2["E"], st["E"]) and np.array_equal(r2["Ae"], st["Ae"])


def test_conjugate_step_moments(oracle):
    """Known answer from conjugacy: one step-2 draw of P has mean (Ap+1+Zin)/(Bp+W E')."""
    M, W, _, _ = synth.cohort(96, 60, 3, burden=3000.0, seed=8)
    st = synth.init_state(96, 60, 3, seed=9)
    Z = np.zeros((96, 60, 3), order="F"); Z[:, :, 1] = np.trunc(M)
    zs = []
    for seed in range(30):
        r = oracle.gibbs_run(M, W, Z, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], burn=0, eval=1, seed=seed,
                             forced_order=[2, 0, 0, 0, 0, 0, 0])
        shape = st["Ap"] + 1 + Z.sum(axis=1); rate = st["Bp"] + W @ st["E"].T
        zs.append((r["P"] - shape / rate) / (np.sqrt(shape) / rate))
    zs = np.array(zs)
    assert abs(zs.mean()) < 5 / math.sqrt(zs.size) and abs(zs.var() - 1) < 0.05


def _prior_draws(rng, n, I, J, K, h):
    """theta ~ prior of the model (R/cpp_eBayesNMF.R:40-50)"""
    Ap = rng.exponential(1 / h["lp"], size=(n, I, K)); Bp = rng.gamma(h["ap"], 1 / h["bp"], size=(n, I, K))
    Ae = rng.exponential(1 / h["le"], size=(n, K, J)); Be = rng.gamma(h["ae"], 1 / h["be"], size=(n, K, J))
    return dict(P=rng.gamma(Ap + 1, 1 / Bp), E=rng.gamma(Ae + 1, 1 / Be), Ap=Ap, Bp=Bp, Ae=Ae, Be=Be)


@pytest.mark.parametrize("be", [2.0, 40.0])        # 40: exposures ~ 20, most counts above 32 (chain + remainder path of step 1)
def test_sweep_leaves_the_joint_distribution_invariant(oracle, be):
    """Exact-invariance test of the WHOLE random-scan sweep (all seven steps, the latent allocation included), in the
    spirit of Geweke (2004): theta ~ prior, M ~ p(M | theta), then a few sweeps given M.  If every step leaves its
    conditional invariant, theta afterwards is again a draw from the prior -- independent replications, so the
    two-sample KS tests against fresh prior draws are exact.  (A 15 % error in one hyper-parameter of the sampler
    drives these p-values below 1e-9.)"""
    I, J, K, R = 2, 3, 2, 8000
    h = dict(ap=3.0, bp=2.0, ae=3.0, be=be, lp=2.0, le=2.0, var_ap=0.5, var_ae=0.5)
    rng = np.random.default_rng(11)
    W = np.asfortranarray(rng.uniform(0.5, 1.5, size=(I, J)))
    th = _prior_draws(rng, R, I, J, K, h)
    big = 0
    after = {k: np.empty_like(v) for k, v in th.items()}
    for r in range(R):
        st = {k: np.asfortranarray(v[r]) for k, v in th.items()}
        M = rng.poisson((st["P"] @ st["E"]) * W).astype(np.float64)
        big += int((M > 32).sum())
        out = oracle.gibbs_run(M, W, None, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], hyper=h, burn=3, eval=0,
                               seed=11000003 + r, want_Z=False)
        for k in after:
            after[k][r] = out[k]
    assert be < 10 or big > R                                          # the large-count path was exercised
    ref = _prior_draws(rng, 4 * R, I, J, K, h)
    pvals = {}
    for k in th:
        a, q = after[k].reshape(R, -1), ref[k].reshape(4 * R, -1)
        pvals[k] = min(stats.ks_2samp(a[:, c], q[:, c]).pvalue for c in range(a.shape[1]))
    lam_a = np.einsum("rik,rkj->rij", after["P"], after["E"]).reshape(R, -1)          # a functional that couples the families
    lam_q = np.einsum("rik,rkj->rij", ref["P"], ref["E"]).reshape(4 * R, -1)
    pvals["PE"] = min(stats.ks_2samp(lam_a[:, c], lam_q[:, c]).pvalue for c in range(lam_a.shape[1]))
    assert min(pvals.values()) > 1e-4, pvals
    assert (after["Ap"] != th["Ap"]).mean() > 0.3 and (after["P"] != th["P"]).all()       # the chain did move


@pytest.mark.parametrize("step", [6, 7])
def test_mh_step_is_stationary_for_its_conditional(oracle, step):
    """Steps 6/7 (src/gibbs_2.cpp:172-251) on one element: the target is p(A | X, B) ~ exp(-l A) (B X)^A / Gamma(A+1).
    Start from exact draws of it (inverse cdf on a fine grid), apply three MH steps, compare with the exact cdf."""
    x, b, l, var, R = 0.37, 2.3, 1.7, 0.6, 6000
    grid = np.linspace(0.0, 14.0, 280001)[1:]
    dens = np.exp(grid * (math.log(b * x) - l) - special.gammaln(grid + 1))
    cdf = np.cumsum(dens); cdf /= cdf[-1]
    rng = np.random.default_rng(step)
    a0 = np.interp(rng.uniform(size=R), cdf, grid)
    hyper = dict(lp=l, le=l, var_ap=var, var_ae=var)
    M, W, one = np.array([[3.0]]), np.array([[1.0]]), np.ones((1, 1))
    out = np.empty(R)
    for r in range(R):
        A = np.array([[a0[r]]])
        st = dict(P=x * one, E=2 * one, Ap=A, Bp=b * one, Ae=one, Be=one) if step == 6 else dict(P=2 * one, E=x * one, Ap=one, Bp=one, Ae=A, Be=b * one)
        res = oracle.gibbs_run(M, W, None, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], hyper=hyper, burn=3, eval=0,
                               seed=77000 + r, forced_order=[[step, 0, 0, 0, 0, 0, 0]] * 3, want_Z=False)
        out[r] = res["Ap" if step == 6 else "Ae"][0, 0]
    assert (out != a0).mean() > 0.3                                    # proposals are accepted
    assert stats.kstest(out, lambda t: np.interp(t, grid, cdf)).pvalue > 1e-3


def test_host_mirror_pieces_cpu(sg):
    from signer_b200.signexp import SignExp, _fivenum_whiskers
    from signer_b200.estimate import EstimateParameters
    assert np.allclose(_fivenum_whiskers(np.array([[1., 2, 3, 4, 5, 6, 7, 8, 9, 100]])), [[1, 3, 5.5, 8, 9]])
    rng = np.random.default_rng(0)
    a, b, l = EstimateParameters(rng.exponential(1 / 1.436, 100000), rng.gamma(2.0506, 1 / 1.325, 100000), 1, 1)
    assert abs(a - 2.0506) < 0.05 and abs(b - 1.325) < 0.04 and abs(l - 1.436) < 0.03
    Ps = rng.gamma(2, 1, (10, 3, 7)); Es = rng.gamma(2, 1, (3, 5, 7))
    SE = SignExp(Ps, Es)
    assert np.allclose(SE.Sign.sum(axis=0), 1.0)                                # R/Definition_SignExp_object.R:120
    assert np.allclose(np.einsum("ikr,kjr->ijr", SE.Sign, SE.Exp), np.einsum("ikr,kjr->ijr", Ps, Es))
    assert np.allclose(SE.median_exp(), np.median(SE.Exp, axis=2))
    # the selection rule on a synthetic BIC curve peaked at 5, ties to the smaller n
    tf = lambda n: [np.full(9, -(n - 5) ** 2 + 0.0), None]
    assert sg.Optimal_sigs(tf, 2, 10, 2)[0] == 5 and sg.Optimal_sigs(tf, 2, 40, 8)[0] == 5
    flat = lambda n: [np.zeros(9), None]
    assert sg.Optimal_sigs(flat, 2, 10, 2)[0] == 2


def test_det_tables_are_the_generators_output_and_identical_on_both_sides(tmp_path):
    """det_tables.h (log table, ziggurat layers): the oracle's and the CUDA library's copies are byte-identical and
    equal to what scripts/make_det_tables.py writes"""
    import filecmp, shutil, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    a, b = os.path.join(root, "oracle", "det_tables.h"), os.path.join(root, "signer_b200", "csrc", "det_tables.h")
    assert filecmp.cmp(a, b, shallow=False)
    work = tmp_path / "repo"
    (work / "scripts").mkdir(parents=True); (work / "oracle").mkdir(); (work / "signer_b200" / "csrc").mkdir(parents=True)
    shutil.copy(os.path.join(root, "scripts", "make_det_tables.py"), work / "scripts" / "make_det_tables.py")
    subprocess.check_call([sys.executable, str(work / "scripts" / "make_det_tables.py")], stdout=subprocess.DEVNULL)
    assert filecmp.cmp(a, str(work / "oracle" / "det_tables.h"), shallow=False)
