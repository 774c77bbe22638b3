"""GPU statistical parity: posterior summaries of the CUDA path against the independent NumPy
oracle (library samplers) and against the planted truth, with BASELINE.json's tolerances written
out: Hungarian-matched signature cosine >= 0.99, median exposure relative error <= 5 %, same
BIC-optimal nsig.  Also the host mirror (eBayesNMF / signeR / Optimal_sigs) end to end."""
import numpy as np
import pytest

from signer_b200 import synth

pytestmark = pytest.mark.gpu

COS_MIN = 0.99
EXPO_REL_ERR_MAX = 0.05


def _report(**kw):
    """numbers worth reading after a GPU run (gpurun_out/ comes back from the box)"""
    import json, os
    print("statistical report:", json.dumps(kw))
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "statistical_report.jsonl"), "a") as f:
            f.write(json.dumps(kw) + "\n")
    except OSError:
        pass


def _planted(I=96, J=240, K=4, seed=5):
    M, W, Pt, Et = synth.cohort(I, J, K, burden=4000.0, seed=seed)
    return M, W, Pt, Et


def test_posterior_matches_numpy_oracle(sg):
    from oracle import numpy_gibbs as NG
    M, W, Pt, Et = _planted()
    I, J = M.shape
    K = 4
    st = synth.init_state(I, J, K, seed=11)
    h = synth.hyper_tuple()
    burn, ev = 600, 400
    got = sg.GibbsSamplerCpp(M, W, None, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], *h, burn, ev,
                             False, False, False, False, False, seed=77)
    ref = NG.gibbs_run(M, W, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], h, burn, ev, seed=78)
    Pg, Eg = NG.summaries(got[2], got[3])
    Pr, Er = NG.summaries(ref["Ps"], ref["Es"])
    perm, cos = NG.match_signatures(Pg, Pr)
    assert cos.min() >= COS_MIN, cos
    # exposures: BASELINE.json's tolerance, median relative error <= 5 %, over ALL K x J entries (no filter) ...
    Er_m = Er[perm, :]
    rel_all = np.abs(Eg - Er_m) / Eg
    med_all = float(np.median(rel_all))
    # ... and on the entries that carry signal (>= 5 % of the genome's total), where Monte Carlo noise of the two
    # 400-sample medians matters least
    big = Eg >= 0.05 * Eg.sum(axis=0, keepdims=True)
    med_big = float(np.median(rel_all[big]))
    _report(exposure_median_rel_err_unfiltered=med_all, exposure_median_rel_err_signal_entries=med_big,
            signature_cosine_min=float(cos.min()))
    assert med_all <= EXPO_REL_ERR_MAX, med_all
    assert med_big <= EXPO_REL_ERR_MAX, med_big
    # both recover the planted signatures
    _, cos_t = NG.match_signatures(Pg, Pt)
    assert cos_t.min() >= 0.97, cos_t
    # the two samplers see the same log-likelihood level (posterior mean within Monte Carlo noise)
    assert abs(np.median(got.bic) - np.median(ref["bic"])) <= 6.0 * (np.std(got.bic) + np.std(ref["bic"]))


def test_gamma_and_binomial_moments_on_device(sg):
    """Conjugacy known answer: with P, Z statistics fixed, one step-3 draw of E has mean
    (Ae+1+Znj)/(Be+P'W) -- checked over many elements at once (z-score of the standardised mean)."""
    M, W, _, _ = _planted(J=400)
    I, J = M.shape
    K = 3
    st = synth.init_state(I, J, K, seed=2)
    h = synth.hyper_tuple()
    Z = np.zeros((I, J, K), order="F"); Z[:, :, 0] = np.trunc(M)
    got = sg.GibbsSamplerCpp(M, W, Z, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], *h, 0, 1,
                             False, False, False, False, True, seed=5, forced_order=[3, 0, 0, 0, 0, 0, 0])
    shape = st["Ae"] + 1 + Z.sum(axis=0).T
    rate = st["Be"] + st["P"].T @ W
    zscore = (got.state["E"] - shape / rate) / (np.sqrt(shape) / rate)
    assert abs(zscore.mean()) < 5.0 / np.sqrt(zscore.size)
    assert abs(zscore.var() - 1.0) < 0.15


def test_ebayesnmf_and_model_selection_breast21(sg):
    """Config 1 (the vignette data): the host mirror end to end with the reference's defaults -- set.seed(42)
    (vignettes/signeR-vignette.Rhtml:78), start = 'lee', testing_burn = testing_eval = 1000 (R/signeR.R:66).  The
    reference's only result-level statement is a figure caption: 'the optimal number of signatures for this dataset
    is 5' (vignettes/signeR-vignette.Rhtml:400).  Single chains of this posterior get trapped in local modes (n = 4, 5
    and 6 lie within ~50 BIC units of each other across seeds), so besides the seed-42 run the test takes the best
    median BIC per candidate over four seeds: both must name 5."""
    M, W, _ = synth.breast21()
    runs = {}
    for seed in (42, 1, 2, 3):
        sg.set_seed(seed)
        runs[seed] = sg.signeR(M.T, samples="rows", Opport=W, nlim=(2, 8), try_all=True, start="lee", testing_burn=1000,
                               testing_eval=1000, main_burn=500, main_eval=200, rng=np.random.default_rng(seed))
    res = runs[42]
    assert res["tested_n"] == [2, 3, 4, 5, 6, 7, 8]
    med = [float(np.median(b)) for b in res["Test_BICs"]]
    nopt = res["Nsign"]
    assert nopt == res["tested_n"][int(np.argmax(med))]           # plain argmax of the median BIC agrees with the replayed rule
    best = np.max([[float(np.median(b)) for b in r["Test_BICs"]] for r in runs.values()], axis=0)
    _report(breast21_nopt_seed42=int(nopt), breast21_median_bics_seed42=med, breast21_nopt_by_seed={str(k): int(v["Nsign"]) for k, v in runs.items()},
            breast21_best_median_bic_per_n=best.tolist())
    assert nopt == 5, (nopt, med)
    assert res["tested_n"][int(np.argmax(best))] == 5, best
    assert all(3 <= r["Nsign"] <= 7 for r in runs.values())
    assert med[res["tested_n"].index(nopt)] > med[0]
    Phat, Ehat = res["Phat"], res["Ehat"]
    assert Phat.shape == (96, nopt) and Ehat.shape == (nopt, 21)
    np.testing.assert_allclose(Phat.sum(axis=0), 1.0, rtol=0.05)   # medians of normalised signatures
    # reconstruction: (Phat Ehat) o W explains the counts
    lam = (Phat @ Ehat) * W
    assert np.corrcoef(lam.ravel(), M.ravel())[0, 1] > 0.98
    SE = res["SignExposures"]
    assert SE.Sign.shape == (96, nopt, 200) and SE.Exp.shape == (nopt, 21, 200)
    # signatures are ordered by total exposure (R/cpp_eBayesNMF.R:190-198)
    tot = Ehat.sum(axis=1) * Phat.sum(axis=0)
    assert np.all(np.diff(tot) <= 1e-9 * tot.max())


def test_em_hyperparameter_path_and_fixed_p(sg):
    M, W, cosmic = synth.breast21()
    rng = np.random.default_rng(3)
    out = sg.eBayesNMF(M, W, 6, ap=1, bp=1, ae=1, be=1, lp=1, le=1, P=cosmic, fixP=True, burn_it=200, eval_it=100,
                       estimate_hyper=True, EM_lim=6, EM_eval_it=50, keep_param=True, rng=rng)
    hyper = out[8]
    assert len(hyper["ae"]) >= 2 and np.isnan(hyper["ap"][0])      # apvet is NA when P is fixed (:91-92)
    assert np.all(np.isfinite(hyper["ae"])) and np.all(hyper["be"] > 0)
    Phat = out[0]
    c = (Phat / np.linalg.norm(Phat, axis=0)).T @ (cosmic / np.linalg.norm(cosmic, axis=0))
    assert np.allclose(np.diag(c), 1.0, atol=1e-9)                 # P fixed: signatures are the given ones
    assert out[4].shape == (96, 6, 100) and out[7].shape == (6, 21, 100)


def test_synthetic_model_selection_recovers_planted_nsig(sg):
    """Config 2 in small: the BIC-optimal nsig of the CUDA path equals the NumPy oracle's and the planted one."""
    from oracle import numpy_gibbs as NG
    M, W, Pt, Et = synth.cohort(96, 160, 3, burden=3000.0, seed=21)
    rng = np.random.default_rng(0)
    res = sg.signeR(M, samples="cols", Opport=W, nlim=(2, 5), try_all=True, start="random",
                    testing_burn=400, testing_eval=200, main_burn=50, main_eval=20, rng=rng)
    med_gpu = {n: float(np.median(b)) for n, b in zip(res["tested_n"], res["Test_BICs"])}
    h0 = synth.hyper_tuple()
    med_ref = {}
    for n in (2, 3, 4, 5):
        h = list(h0); h[3] = h0[3] / np.sqrt(n)                     # be/sqrt(n), R/signeR.R:163
        st = synth.init_state(96, 160, n, hyper=dict(be=h[3]), seed=100 + n)
        r = NG.gibbs_run(M, W, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], tuple(h), 400, 200, seed=n)
        med_ref[n] = float(np.median(r["bic"]))
    assert max(med_gpu, key=med_gpu.get) == max(med_ref, key=med_ref.get) == 3, (med_gpu, med_ref)
    assert res["Nsign"] == 3


@pytest.mark.parametrize("R", [7, 64, 301])
def test_device_summaries_equal_host_medians(sg, R):
    """Phat/Ehat computed on the device (normalisation + medians over the eval samples) against NumPy on the same samples."""
    M, W, _, _ = synth.cohort(24, 50, 3, burden=500.0, seed=1)
    st = synth.init_state(24, 50, 3, seed=2)
    with sg.Chain(M, W, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], synth.hyper_tuple(), seed=3) as c:
        res = c.run(burn=20, eval=R, keep_par=False, fetch=True, summaries=True)
    Ps, Es = res[2], res[3]
    s = np.zeros((3, R))
    for i in range(24):                                   # sequential sums like the device
        s = s + Ps[i, :, :]
    Phat = np.median(Ps / s[None, :, :], axis=2)
    Ehat = np.median(Es * s[:, None, :], axis=2)
    np.testing.assert_allclose(res.Phat, Phat, rtol=1e-15, atol=0)
    np.testing.assert_allclose(res.Ehat, Ehat, rtol=1e-15, atol=0)
    assert np.allclose(res.Phat.sum(axis=0), 1.0, atol=0.05)


def test_ebayesnmf_device_summaries_match_signexp_path(sg):
    M, W, _ = synth.breast21()
    a = sg.eBayesNMF(M, W, 4, *[synth.HYPER_DEFAULT[k] for k in ("ap", "bp", "ae", "be", "lp", "le")], burn_it=100, eval_it=60,
                     rng=np.random.default_rng(5))
    b = sg.eBayesNMF(M, W, 4, *[synth.HYPER_DEFAULT[k] for k in ("ap", "bp", "ae", "be", "lp", "le")], burn_it=100, eval_it=60,
                     rng=np.random.default_rng(5), keep_samples=False, device_summaries=True)
    np.testing.assert_allclose(b[0], a[0], rtol=1e-13)
    np.testing.assert_allclose(b[1], a[1], rtol=1e-13)
    assert b[3] is None and np.array_equal(a[2], b[2])
