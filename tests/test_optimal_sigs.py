"""Optimal_sigs (signer_b200/signer.py) against the rule of R/Optimal_sigs.R:1-168, on BIC tables where the literal rule
and a 'compare everything that was evaluated' rule disagree.  CPU only: testfun is a table look-up."""
import numpy as np

import signer_b200 as sg


def _table(med):
    calls = []

    def tf(n):
        calls.append(n)
        return [np.full(5, float(med[n])) + np.arange(5) * 1e-6, dict(n=n)]
    return tf, calls


def test_unimodal_tables_find_the_peak_for_any_worker_count():
    med = {n: -abs(n - 7) for n in range(1, 41)}
    for ncores in (1, 2, 3, 8, 64):
        tf, _ = _table(med)
        nopt, tested, res = sg.Optimal_sigs(tf, 1, 40, 8, ncores=ncores)
        assert nopt == 7, (ncores, nopt, tested)
        assert tested == sorted(tested) and len(res) == len(tested) and 7 in tested


def test_downhill_exit_never_compares_what_was_still_running():
    """Coarse grid 1,5,9,...,29,32 with three workers.  First poll: 1, 5, 9 are compared and 13, 17, 21 are launched; two
    evaluated candidates lie above max(nmax) = 1, so the coarse loop ends (R/Optimal_sigs.R:29,80-83).  13, 17, 21 finish
    and are reported (:92-98) but never compared; 25, 29, 32 were never launched (:99-105 overwrites eval_latter).
    The table's second, higher peak at 21 is therefore invisible to the reference's rule."""
    med = {n: -float(n) for n in range(1, 33)}
    med[21] = 100.0                                         # a much better candidate among the stragglers
    tf, calls = _table(med)
    nopt, tested, _ = sg.Optimal_sigs(tf, 1, 32, 4, ncores=3)
    assert nopt == 1
    assert set([13, 17, 21]) <= set(tested)                 # finished and reported ...
    assert not ({25, 29, 32} & set(calls))                  # ... the rest never ran
    # a rule that compares every evaluated candidate would have moved to 21
    best_of_all = max(tested, key=lambda n: med[n])
    assert best_of_all == 21 and best_of_all != nopt
    # with enough workers every coarse candidate is polled in the first pass and 21 wins
    tf, _ = _table(med)
    assert sg.Optimal_sigs(tf, 1, 32, 4, ncores=16)[0] == 21


def test_refinement_single_element_quirk():
    """R/Optimal_sigs.R:98-100: when more candidates are left than workers are free the reference launches
    to_eval[free] -- ONE element -- and queues to_eval[-(1:free)]; elements before it are never evaluated."""
    med = {n: -abs(n - 9) for n in range(1, 18)}
    tf, calls = _table(med)
    nopt, tested, _ = sg.Optimal_sigs(tf, 1, 17, 8, ncores=2)
    # coarse: 1, 9 compared (nmax = 9), 17 launched and compared in the second pass; no downhill (only one above 9)
    # refine step 4: to_eval = [5, 13], two free workers -> both; step 2: [7, 11]; step 1: [8, 10]
    assert nopt == 9 and set(calls) == {1, 9, 17, 5, 13, 7, 11, 8, 10}
    tf, calls = _table(med)
    nopt1, tested1, _ = sg.Optimal_sigs(tf, 1, 17, 8, ncores=1)
    # one worker: each refinement launches only to_eval[1] and queues the rest minus the first (:100): 5 then 13, ...
    assert nopt1 == 9 and tested1 == sorted(set(calls))


def test_ties_prefer_fewer_signatures_and_significance_mode():
    flat = {n: 0.0 for n in range(2, 11)}

    def tf(n):
        return [np.zeros(7) + flat[n], None]
    assert sg.Optimal_sigs(tf, 2, 10, 2)[0] == 2            # R/Optimal_sigs.R:159
    rng = np.random.default_rng(3)

    def noisy(n):
        return [rng.normal(-abs(n - 5) * 50.0, 1.0, size=40), None]
    assert sg.Optimal_sigs(noisy, 2, 10, 2, significance=True, pcrit=0.05, ncores=8)[0] == 5
