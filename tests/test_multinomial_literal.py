"""The draw specification's step 1 (classes visited heaviest first, chain stopped at 32 left, direct method for small
counts and for the remainder) against R's rmultinom structure taken LITERALLY (classes in index order, conditional
binomials to the end: oracle_multinomial_cell_literal).  The two are different algorithms for the
same law, so they are compared statistically: every class marginal against its exact Binomial(n, p_k) moments and the
two empirical distributions against each other (two-sample chi-square per class) -- on the CPU for the oracle's
restatement of the specification, and on the GPU for the kernel itself."""
import numpy as np
import pytest
from scipy import stats

CASES = [  # (count, class weights)
    (7, [0.5, 0.2, 0.2, 0.1]),
    (32, [0.05, 0.6, 0.05, 0.3]),
    (33, [0.3, 0.3, 0.3, 0.1]),
    (180, [0.02, 0.5, 0.08, 0.25, 0.15]),
    (2500, [0.001, 0.7, 0.05, 0.2, 0.049]),
    (40000, [0.25, 0.25, 0.0, 0.25, 0.25]),
]
REPS = 4000


def _check(z_spec, z_lit, n, w):
    p = np.asarray(w) / np.sum(w)
    assert np.all(z_spec.sum(axis=1) == n) and np.all(z_lit.sum(axis=1) == n)
    for k in range(len(w)):
        for name, z in (("specification", z_spec), ("literal", z_lit)):
            m, v = z[:, k].mean(), z[:, k].var()
            se = np.sqrt(n * p[k] * (1 - p[k]) / len(z)) + 1e-12
            assert abs(m - n * p[k]) <= 5.0 * se, (name, n, k, m, n * p[k])
            if p[k] > 0:
                assert abs(v - n * p[k] * (1 - p[k])) <= 0.12 * n * p[k] * (1 - p[k]) + 0.05, (name, n, k, v)
        if p[k] == 0:
            assert z_spec[:, k].max() == 0 and z_lit[:, k].max() == 0
            continue
        # two-sample chi-square on the pooled support, bins with at least ~10 expected
        lo, hi = int(min(z_spec[:, k].min(), z_lit[:, k].min())), int(max(z_spec[:, k].max(), z_lit[:, k].max()))
        edges = np.unique(np.quantile(np.concatenate([z_spec[:, k], z_lit[:, k]]), np.linspace(0, 1, 13)))
        if len(edges) < 3:
            continue
        edges[0], edges[-1] = lo - 0.5, hi + 0.5
        a, _ = np.histogram(z_spec[:, k], bins=edges)
        b, _ = np.histogram(z_lit[:, k], bins=edges)
        keep = (a + b) >= 20
        if keep.sum() < 2:
            continue
        chi2, pval, _, _ = stats.chi2_contingency(np.vstack([a[keep], b[keep]]))
        assert pval > 1e-5, (n, k, pval)


@pytest.mark.parametrize("n,w", CASES)
def test_specification_and_literal_rmultinom_have_the_same_law(oracle, n, w):
    K = len(w)
    P = np.asfortranarray(np.array([w], dtype=np.float64))          # one context
    E = np.asfortranarray(np.ones((K, 1)) * 3.7)
    z_spec = np.array([oracle.multinomial_cell(1000 + r, 0, 1, 0, 0, n, P, E) for r in range(REPS)])
    z_lit = np.array([oracle.multinomial_cell_literal(5000 + r, 0, 1, 0, 0, n, P, E) for r in range(REPS)])
    _check(z_spec, z_lit, n, w)


@pytest.mark.gpu
@pytest.mark.parametrize("n,w", CASES)
def test_cuda_step1_and_literal_rmultinom_have_the_same_law(sg, oracle, n, w):
    """One launch of z_alloc_fused on a cohort whose cells are i.i.d. replicates of one (count, probabilities) pair: the
    final Z cube of a step-1-only sweep gives REPS draws of the kernel's multinomial."""
    from signer_b200 import synth
    K = len(w)
    I, J = 40, REPS // 40
    P = np.asfortranarray(np.tile(np.array(w, dtype=np.float64), (I, 1)))
    E = np.asfortranarray(np.ones((K, J)) * 3.7)
    M = np.asfortranarray(np.full((I, J), float(n)))
    W = np.asfortranarray(np.ones((I, J)))
    one = np.asfortranarray(np.ones((I, K))); onee = np.asfortranarray(np.ones((K, J)))
    got = sg.GibbsSamplerCpp(M, W, None, P, E, one, one, onee, onee, *synth.hyper_tuple(), 0, 1, False, False, False, False, True,
                             seed=77, forced_order=[1, 0, 0, 0, 0, 0, 0])
    z_gpu = got[0][:, :, :, 0].reshape(I * J, K)
    z_lit = np.array([oracle.multinomial_cell_literal(9000 + r, 0, 1, 0, 0, n, P[:1], E[:, :1]) for r in range(I * J)])
    _check(z_gpu, z_lit, n, w)
