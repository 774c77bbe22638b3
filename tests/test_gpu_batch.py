"""Candidates x chains through signer_gibbs_batch (all visible GPUs): every task must give exactly
what a stand-alone run with the same state and seed gives."""
import numpy as np
import pytest

from helpers import assert_bits_equal
from signer_b200 import _lib, synth
from signer_b200.signer import batch_bics

pytestmark = pytest.mark.gpu


def test_batch_equals_single_runs(sg):
    M, W, _, _ = synth.cohort(96, 130, 4, burden=1500.0, seed=4)
    nd = _lib.lib().signer_device_count()
    tasks = []
    for n in (2, 3, 5, 7, 4, 6):
        h = list(synth.hyper_tuple()); h[3] = h[3] / np.sqrt(n)
        tasks.append(dict(K=n, state=synth.init_state(96, 130, n, seed=n), hyper=tuple(h), burn=30, eval=12, seed=900 + n))
    res = batch_bics(M, W, tasks, devices=list(range(nd)))
    for tk, (ll, bic) in zip(tasks, res):
        st = tk["state"]
        one = sg.GibbsSamplerCpp(M, W, None, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], *tk["hyper"], 30, 12,
                                 False, False, False, False, False, seed=tk["seed"])
        assert_bits_equal(ll, one.loglik, "loglik K=%d" % tk["K"])
        assert_bits_equal(bic, one.bic, "bic K=%d" % tk["K"])


def test_batch_reports_bad_task(sg):
    M, W, _, _ = synth.cohort(20, 30, 2, seed=1)
    st = synth.init_state(20, 30, 200, seed=1)       # K = 200 is above the supported maximum
    with pytest.raises(sg.SignerError):
        batch_bics(M, W, [dict(K=200, state=st, hyper=synth.hyper_tuple(), burn=1, eval=1, seed=1)], devices=[0])
