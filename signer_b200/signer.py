"""signeR() and Optimal_sigs(): the public entry and the nsig model-selection loop, mirroring
R/signeR.R:62-230 and R/Optimal_sigs.R:1-168 on top of the CUDA sampler.

The reference evaluates candidates as `future`s on availableCores()-1 PSOCK worker processes and
polls them.  Optimal_sigs below is a statement-by-statement transcription of R/Optimal_sigs.R; the
only substitution is the future itself: a launched candidate is queued, and the first poll of a
queued candidate evaluates everything queued in one batch call (one host thread per GPU,
signer_gibbs_batch) -- the reference's behaviour when its workers finish between two polls.
"""
import ctypes as C
import math

import numpy as np

from . import _lib, gibbs
from .ebayes import eBayesNMF, initial_state

HH_DEFAULT = dict(ap=0.8005, bp=0.043, ae=2.0506, be=1.325, lp=5.3329, le=1.436)   # R/signeR.R:110


def _compare(thisbics, nmax, values, significance, pcrit):
    """The 'compare k and nmax' block of R/Optimal_sigs.R:38-62 (and :112-137): one verdict per member of nmax,
    1 = k is better, -1 = k is worse, 0 = tie (or not significantly different)."""
    thismedian = np.median(thisbics)
    testeq = []
    for n in nmax:
        maxbics = values[n][0]
        maxmed = np.median(maxbics)
        if significance:
            from scipy import stats
            p = stats.kruskal(thisbics, maxbics).pvalue
            testeq.append((1 if thismedian > maxmed else -1) if p <= pcrit else 0)
        else:
            testeq.append(1 if thismedian > maxmed else (-1 if thismedian < maxmed else 0))
    return testeq


def _update_nmax(k, thisbics, nmax, values, significance, pcrit):
    """R/Optimal_sigs.R:63-65 (and :138-139): keep those n not worse than k; keep k if no other n is better."""
    testeq = _compare(thisbics, nmax, values, significance, pcrit)
    nmax = [n for n, t in zip(nmax, testeq) if t in (0, -1)]
    if len(nmax) == 0 or min(testeq) >= 0:
        nmax.append(k)
    return nmax


def Optimal_sigs(testfun, liminf, limsup, step, significance=False, pcrit=0.05, parplan="multisession", *,
                 batchfun=None, ncores=None):
    """Literal transcription of R/Optimal_sigs.R:1-168, statement by statement (line cites on the right).

    testfun(n) -> [bics, HH].  The reference launches testfun(n) as `future`s on worker processes and polls
    `resolved()`; here `launch` queues a candidate and the first poll of a queued candidate evaluates EVERYTHING
    queued so far in one call of batchfun(list_of_n) -> {n: [bics, HH]} (one multi-GPU batch, signer_gibbs_batch).
    That is the reference's behaviour when its workers finish between two polls; like there, which candidates get
    evaluated and compared depends on `ncores` (parallel::detectCores()-1, :17): candidates still running when the
    coarse search turns downhill are finished but never compared (:29, :92-98), candidates not yet launched at that
    point are dropped (eval_latter is overwritten, :99-105), and the refinement takes ONE element of to_eval when
    more are left than workers are free (:100, as written in the reference).
    Returns [nopt, evaluated n (ascending), their results] (:159-167)."""
    import os
    if batchfun is None:
        def batchfun(ns):
            return {n: testfun(n) for n in ns}
    if ncores is None:
        ncores = (os.cpu_count() or 2) - 1                                 # :17
    ncores = max(1, int(ncores))
    values, queued = {}, []

    def launch(k):                                                         # values[[k]] %<-% {testfun(k)}
        queued.append(k)

    def resolved(k):                                                       # resolved(values)[ind]
        if queued:
            values.update(batchfun(list(queued)))
            del queued[:]
        return k in values

    to_eval = list(range(liminf, limsup + 1, step))                        # :14
    if to_eval[-1] < limsup:
        to_eval.append(limsup)                                             # :15
    evalued, nmax, downhill = [], [], False                                # :16-19
    if len(to_eval) > ncores:                                              # :20-26
        eval_now, eval_latter = to_eval[:ncores], to_eval[ncores:]
    else:
        eval_now, eval_latter = list(to_eval), []
    for k in eval_now:                                                     # :27
        launch(k)
    running = list(eval_now)                                               # :28
    while not downhill and len(running) > 0:                               # :29 large steps, looking for downhill descent
        still_running = list(running)                                      # :30
        for k in still_running:                                            # :31
            if resolved(k):                                                # :33
                thisbics = values[k][0]                                    # :34
                if len(nmax) > 0:                                          # :37
                    nmax = _update_nmax(k, thisbics, nmax, values, significance, pcrit)   # :38-65
                else:
                    nmax = [k]                                             # :67
                running = [r for r in running if r != k]                   # :69
                evalued.append(k)                                          # :70
                if len(eval_latter) > 0:                                   # :71-77 start testing another number of signatures
                    nextrun = eval_latter.pop(0)
                    launch(nextrun)
                    running.append(nextrun)
        if len(nmax) > 0 and len(evalued) >= 3:                            # :80-83 check downhill
            if sum(1 for e in evalued if e > max(nmax)) >= 2:
                downhill = True
    while step > 1:                                                        # :85 refining steps
        step = int(math.ceil(step / 2.0))                                  # :86
        to_eval = sorted(set([n - step for n in nmax] + [n + step for n in nmax]))        # :87
        to_eval = [t for t in to_eval if liminf <= t <= limsup]            # :88
        to_eval = [t for t in to_eval if t not in evalued and t not in running]           # :89
        left_eval = len(to_eval)                                           # :90
        for k in list(running):                                            # :91-97 finished meanwhile: reported, NOT compared
            if resolved(k):
                running = [r for r in running if r != k]
                evalued.append(k)
        free = ncores - len(running)
        if left_eval > free:                                               # :98-100, R's 1-based single-element index as written
            eval_now = [to_eval[free - 1]] if free >= 1 else []
            eval_latter = to_eval[free:] if free >= 1 else to_eval[1:]     # to_eval[-c(1:free)]; 1:0 is c(1,0): drops the first
        else:                                                              # :101-104
            eval_now, eval_latter = list(to_eval), []
        for k in eval_now:                                                 # :106
            launch(k)
        running = running + eval_now                                       # :107
        while len(eval_now) > 0:                                           # :108
            still_running = list(eval_now)                                 # :109
            for k in still_running:                                        # :110
                if resolved(k):                                            # :112
                    thisbics = values[k][0]
                    nmax = _update_nmax(k, thisbics, nmax, values, significance, pcrit)   # :115-139
                    eval_now = [e for e in eval_now if e != k]             # :141
                    running = [r for r in running if r != k]               # :142
                    evalued.append(k)                                      # :143
                    if len(eval_latter) > 0:                               # :144-150
                        nextrun = eval_latter.pop(0)
                        launch(nextrun)
                        eval_now.append(nextrun)
    nopt = min(nmax)                                                       # :159 in case of ties, less signatures are better
    for k in running:                                                      # :160-165
        if resolved(k):
            evalued.append(k)
    tested = [n for n in range(liminf, limsup + 1) if n in evalued]        # :166
    return [nopt, tested, [values[n] for n in tested]]                     # :167


def batch_bics(M, W, tasks, devices=None):
    """Candidates x chains in ONE call of signer_gibbs_batch (include/signer_gibbs.h).
    tasks: list of dicts(K, state, hyper (8-tuple), burn, eval, seed).  Returns a list of
    (loglik, bic) arrays in task order."""
    L = _lib.lib()
    M = np.asfortranarray(np.asarray(M, dtype=np.float64))
    W = np.asfortranarray(np.asarray(W, dtype=np.float64))
    I, J = M.shape
    if devices is None:
        nd = L.signer_device_count()
        if nd < 1:
            raise _lib.SignerError(3, "no CUDA device (there is no CPU fallback)")
        devices = list(range(nd))
    arr = (_lib.Task * len(tasks))()
    keep = []
    for t, tk in enumerate(tasks):
        st = {k: np.asfortranarray(np.asarray(v, dtype=np.float64)) for k, v in tk["state"].items()}
        ll = np.zeros(tk["eval"]); bic = np.zeros(tk["eval"])
        keep.append((st, ll, bic))
        arr[t].K = int(tk["K"])
        arr[t].state = _lib.State(None, *[st[k].ctypes.data_as(_lib.dp) for k in ("P", "E", "Ap", "Bp", "Ae", "Be")])
        arr[t].opts = gibbs._make_opts(tk["hyper"], tk["burn"], tk["eval"], False, False, False, False, False,
                                       tk["seed"], 0, 0, None)
        arr[t].loglik = ll.ctypes.data_as(_lib.dp)
        arr[t].bic = bic.ctypes.data_as(_lib.dp)
    dev = np.asarray(devices, dtype=np.int32)
    rc = L.signer_gibbs_batch(I, J, M.ctypes.data_as(_lib.dp), W.ctypes.data_as(_lib.dp), arr, len(tasks),
                              dev.ctypes.data_as(_lib.ip), len(devices))
    _lib.check(rc)
    return [(k[1], k[2]) for k in keep]


def signeR(M, Mheader=True, samples="rows", Opport=None, Oppheader=False, P=None, fixedP=False, nsig=None,
           nlim=(None, None), try_all=False, BICsignificance=False, critical_p=0.05,
           ap=None, bp=None, ae=None, be=None, lp=None, le=None, var_ap=10, var_ae=10,
           start="lee", testing_burn=1000, testing_eval=1000, main_burn=10000, main_eval=2000,
           estimate_hyper=False, EMit_lim=100, EM_eval=100, parallelization="multisession",
           *, rng=None, devices=None, verbose=False):
    """Mirror of signeR() (R/signeR.R:62-230; documented in man/signeR.Rd).  M: samples x contexts
    when samples == "rows" (the reference's default), else contexts x samples.  File reading
    (read.table paths) is input preparation and out of scope: pass arrays.  `parallelization` is
    accepted for signature compatibility; candidates are spread over `devices` (all GPUs by default).
    Returns the reference's result list as a dict with the same names."""
    rng = rng if rng is not None else np.random.default_rng(gibbs.next_seed())
    Mread = np.asarray(M, dtype=np.float64)
    if Mread.ndim != 2:
        raise ValueError("M should be a matrix, a data frame or a link to a file.")
    M = Mread.T if samples == "rows" else Mread                            # :78-81
    if Opport is None:
        Wm = np.ones_like(M)                                               # :95-97
    else:
        Opp = np.asarray(Opport, dtype=np.float64)
        if Opp.shape == M.shape:
            Wm = Opp
        elif Opp.T.shape == M.shape:
            Wm = Opp.T
        else:
            raise ValueError("Dimensions of M and Opport do not match.")    # :102-107
    HH = dict(HH_DEFAULT)
    if not estimate_hyper:                                                 # :111-118
        ap = HH["ap"] if ap is None else ap; bp = HH["bp"] if bp is None else bp
        ae = HH["ae"] if ae is None else ae; be = HH["be"] if be is None else be
        lp = HH["lp"] if lp is None else lp; le = HH["le"] if le is None else le
    if P is None:
        fixedP = False
    else:
        P = np.asarray(P, dtype=np.float64)
        nsig = P.shape[1]
    if fixedP:                                                             # :120-129
        if any(v is None for v in (ae, be, le)):
            eh = True
            ae = 1 if ae is None else ae; be = 1 if be is None else be; le = 1 if le is None else le
        else:
            eh = estimate_hyper
    else:                                                                  # :130-142
        if any(v is None for v in (ap, bp, ae, be, lp, le)):
            eh = True
            ap = 1 if ap is None else ap; bp = 1 if bp is None else bp; ae = 1 if ae is None else ae
            be = 1 if be is None else be; lp = 1 if lp is None else lp; le = 1 if le is None else le
        else:
            eh = estimate_hyper
    tested_n, Test_BICs, Hyper_paths = None, None, None
    if nsig is None and not fixedP:                                        # :143-194 search for the best n
        Nmin = 1 if nlim[0] is None else int(nlim[0])
        Nmax = min(M.shape) - 1 if nlim[1] is None else int(nlim[1])
        step0 = 1 if try_all else 2 ** max(int(math.floor(math.log2(Nmax - Nmin + 1))) - 2, 0)   # :154-158
        I, J = M.shape

        def hyper_for(n):
            return (ap, bp, ae, be if eh else be / math.sqrt(n), lp, le)   # :163

        def testfun(n):
            a_, b_, ae_, be_, lp_, le_ = hyper_for(n)
            r = eBayesNMF(M, Wm, n, a_, b_, ae_, be_, lp_, le_, var_ap, var_ae, P=None, fixP=False, start=start,
                          burn_it=testing_burn, eval_it=testing_eval, estimate_hyper=eh, EM_lim=EMit_lim,
                          EM_eval_it=EM_eval, keep_param=False, rng=rng, verbose=verbose, keep_samples=False)
            return [r[2], r[8]]                                            # :169-172 only the BICs and the hyper path

        def batchfun(ns):
            if eh:
                return {n: testfun(n) for n in ns}
            # without EM a candidate is one sampler call: evaluate the whole wave in one multi-GPU batch
            tasks = []
            for n in ns:
                h6 = hyper_for(n)
                st = dict(zip(("P", "E", "Ap", "Bp", "Ae", "Be"),
                              initial_state(np.asfortranarray(M), np.asfortranarray(Wm), n, *h6, None, False, start, rng)))
                tasks.append(dict(K=n, state=st, hyper=h6 + (var_ap, var_ae), burn=testing_burn, eval=testing_eval,
                                  seed=int(rng.integers(0, 2 ** 63 - 1))))
            res = batch_bics(M, Wm, tasks, devices)
            out = {}
            for n, (ll, bic) in zip(ns, res):
                h6 = hyper_for(n)
                out[n] = [bic, dict(zip(("ap", "bp", "ae", "be", "lp", "le"), [np.array([v]) for v in h6]))]
            return out

        Ops = Optimal_sigs(testfun, Nmin, Nmax, step0, significance=BICsignificance, pcrit=critical_p, batchfun=batchfun)
        nopt, tested_n = Ops[0], Ops[1]
        Test_BICs = [r[0] for r in Ops[2]]
        Hyper_paths = [r[1] for r in Ops[2]]
        best = Hyper_paths[tested_n.index(nopt)]                           # :186-193
        ap, bp, ae, be, lp, le = (float(best[k][-1]) for k in ("ap", "bp", "ae", "be", "lp", "le"))
        eh = False
    else:
        nopt = P.shape[1] if fixedP else int(nsig)
    final = eBayesNMF(M, Wm, nopt, ap, bp, ae, be, lp, le, var_ap, var_ae, P=P, fixP=fixedP, start=start,
                      burn_it=main_burn, eval_it=main_eval, estimate_hyper=eh, EM_lim=EMit_lim, EM_eval_it=EM_eval,
                      keep_param=True, rng=rng, device=(devices[0] if devices else 0), verbose=verbose)   # :211-215
    return dict(Nsign=nopt, tested_n=tested_n, Test_BICs=Test_BICs, Phat=final[0], Ehat=final[1],
                SignExposures=final[3], BICs=final[2], Hyper_paths=Hyper_paths)          # :220-227
