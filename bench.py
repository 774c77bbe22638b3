#!/usr/bin/env python
"""bench.py -- Gibbs sweeps/s of the B200 sampler on BASELINE.json's headline workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is ONE full Gibbs sweep (all seven steps of src/gibbs_2.cpp:305-318) of one chain over the
whole cohort.  Workload (config.workload): BASELINE.json configs[2], the configuration the metric
is quoted on: synthetic 96 x 10,000 genomes with an opportunity matrix, K = 20, one long chain per
GPU.  With N > 1 GPUs every rank runs its own independent chain on the same cohort shape (the
path shards by chains/candidates, no data-path collective): weak scaling.

One JSON line on stdout (rank 0).  Keys beyond the base contract:
  value            burn-in sweeps/s, chain resident in HBM, CUDA events on the launching stream
  eval_value       evaluation sweeps/s (samples into the device ring + log-likelihood every sweep)
  e2e              sweeps/s through the reference-shaped call GibbsSamplerCpp(...) with HOST arrays:
                   M, W, the I x J x K cube Z and the state go in, the sample cubes come back
                   (burn = eval = K/2, keep_par = FALSE: the reference's model-selection call,
                   R/signeR.R:164-168), host<->device copies inside the timed region
  kernel_ms, kernel_ms_eval   mean launch time of each kernel in burn-in / evaluation sweeps (second and third pass
                   with CUDA events around every launch)
  roofline         dominant kernel z_alloc_fused: algorithmic bytes per launch / mean launch time
  cpu_baseline     the reference's src/gibbs_2.cpp (oracle/_ref) -- or the oracle port -- on one host core of this box
  em_loop          ten reference-shaped calls of 200 burn-in + 100 evaluation sweeps with keep_par=TRUE, each fed the
                   previous call's last samples (the EM loop of R/cpp_eBayesNMF.R:104-152 as R drives it), host arrays
  nsig             the second half of BASELINE.json's metric: wall time of the model-selection sweep nsig = 2..20 x 4
                   chains, 1000 + 1000 sweeps each (R/signeR.R:66), on the same cohort, over the N ranks (strong scaling)
  sharded          BASELINE.json config 4 (1536 x 20,000, K = 30): ONE chain with its genome columns sharded over the
                   N GPUs, driven by rank 0 (ms per sweep), plus a small sharded chain over the same devices checked
                   bit for bit against the oracle (sharded_parity_ok; NCCL when N > 1, emulated shards when N = 1)
Timing: whatever --warmup says, at least 200 untimed sweeps precede the timed region (a 5-sweep warm-up leaves the
L2 / instruction caches and clocks cold: 19 % slower sweeps); --steps stays the timed count.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG", "NONE")      # keep stdout to the one JSON line (no NCCL version banner)

MIN_WARMUP_SWEEPS = 200          # untimed, whatever --warmup says (see the docstring)
METRIC = "gibbs_sweeps_per_s"
UNIT = "sweeps/s"
WORKLOADS = {
    "cfg3_96x10000_K20": dict(K=20, desc="synthetic 96x10000 genomes (pan-TCGA scale) with opportunity matrix, K=20, single long chain"),
    "cfg2_96x560": dict(K=10, desc="synthetic 96x560 genomes (ICGC-breast scale), K=10"),
    "cfg5_96x100000": dict(K=40, desc="synthetic 96x100000 genomes, K=40"),
    "cfg4_1536x20000_K30": dict(K=30, desc="synthetic 1536x20000, K=30"),
}


def config_dict(workload, I, J, K):
    """identical in both arms (the driver compares them)"""
    return dict(workload=workload, description=WORKLOADS[workload]["desc"], I=I, J=J, K=K, chains_per_gpu=1,
                l2="no flush between steps: a chain iterates over the same cohort every sweep, its working set (M int32 + W f64 "
                   "= 11.5 MB, state 5 MB) is L2-resident by design")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def cpu_sweep_rate(M, W, K, seconds_target, seed=1):
    """Time the CPU oracle on a bounded sample: full sweeps over the first Js genomes (cost is linear
    in genomes), scaled back to the full cohort.  Returns (full-cohort sweeps/s, description)."""
    from oracle import oracle as _O
    from signer_b200 import synth
    kind = "port"
    O = _O
    try:
        from oracle import ref_runner
        if ref_runner.available():       # the reference's own gibbs_2.cpp built against stand-in headers
            O, kind = ref_runner, "reference"
    except Exception:
        pass
    I, J = M.shape
    Js = min(J, 200)
    st = synth.init_state(I, Js, K, seed=seed)
    t0 = time.perf_counter()
    O.gibbs_run(M[:, :Js], W[:, :Js], None, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], burn=1, eval=0,
                seed=seed, want_Z=False)
    per_col = (time.perf_counter() - t0) / 2.0 / Js          # init allocation + 1 sweep
    # choose sample: nsw sweeps over Js columns ~ seconds_target
    nsw = 4
    Js = int(max(32, min(J, seconds_target / (nsw + 1) / per_col)))
    st = synth.init_state(I, Js, K, seed=seed)
    Msub, Wsub = np.asfortranarray(M[:, :Js]), np.asfortranarray(W[:, :Js])
    t0 = time.perf_counter()
    O.gibbs_run(Msub, Wsub, None, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], burn=0, eval=0, seed=seed, want_Z=False)
    t_init = time.perf_counter() - t0
    t0 = time.perf_counter()
    O.gibbs_run(Msub, Wsub, None, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], burn=nsw, eval=0, seed=seed, want_Z=False)
    t_run = time.perf_counter() - t0 - t_init
    per_sweep_full = t_run / nsw * (J / Js)
    what = "the reference's src/gibbs_2.cpp (oracle/_ref)" if kind == "reference" else "the oracle port"
    return 1.0 / per_sweep_full, kind, "%d burn-in sweeps of %s over the first %d of %d genomes, scaled by %d/%d" % (nsw, what, Js, J, J, Js)


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU implementation of the path on this box's host cores.
    The reference itself needs R + Rcpp + RcppArmadillo (absent), so this times oracle/_ref when the
    reference sources could be compiled against the stand-in headers, else the oracle port.  One
    chain = one thread (the reference's native code is single-threaded, R/Optimal_sigs.R:3-8 only
    parallelises over candidates)."""
    if rank != 0:
        return
    from signer_b200 import synth
    M, W, _, _ = synth.named(args.workload)
    K = WORKLOADS[args.workload]["K"]
    I, J = M.shape
    from oracle import oracle as O
    kind = "port"
    runner = O.gibbs_run
    try:
        from oracle import ref_runner
        if ref_runner.available():
            runner = ref_runner.gibbs_run
            kind = "reference"
    except Exception:
        pass
    # size the per-step sample so warmup+steps finish in about two minutes
    Js0 = min(J, 128)
    st = synth.init_state(I, Js0, K, seed=3)
    t0 = time.perf_counter()
    runner(np.asfortranarray(M[:, :Js0]), np.asfortranarray(W[:, :Js0]), None, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"],
           burn=2, eval=0, seed=3, want_Z=False)
    per_col_sweep = (time.perf_counter() - t0) / 3.0 / Js0
    budget = 120.0
    Js = int(max(16, min(J, budget / max(1, args.steps + args.warmup) / per_col_sweep)))
    st = synth.init_state(I, Js, K, seed=3)
    Msub, Wsub = np.asfortranarray(M[:, :Js]), np.asfortranarray(W[:, :Js])
    t0 = time.perf_counter()
    runner(Msub, Wsub, None, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], burn=args.warmup, eval=0, seed=3, want_Z=False)
    t_warm = time.perf_counter() - t0
    t0 = time.perf_counter()
    runner(Msub, Wsub, None, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], burn=args.warmup + args.steps, eval=0, seed=3, want_Z=False)
    t_all = time.perf_counter() - t0
    dt = max(t_all - t_warm, 1e-9)
    ms_step_sample = dt / args.steps * 1e3
    value = args.steps / dt * (Js / J)
    sample = ("each step = one Gibbs sweep over the first %d of %d genomes (cost linear in genomes); "
              "value scaled by %d/%d to the full cohort" % (Js, J, Js, J))
    line = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms_step_sample * (J / Js), higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                data="synthetic", config=config_dict(args.workload, I, J, K),
                cpu_baseline=dict(value=value, unit=UNIT, cores=1, kind=kind, sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


def run_nsig(args):
    """Secondary metric: wall time of the nsig model-selection sweep (every candidate n in the range x
    `chains` chains, testing_burn = testing_eval = 1000 as R/signeR.R:66) on N GPUs, one process per GPU,
    candidates dealt longest-first over the ranks; only BIC vectors travel."""
    import torch
    from signer_b200 import dist, synth
    from signer_b200.signer import batch_bics
    rank, world, local = dist.init()
    wl = args.workload
    M, W, _, _ = synth.named(wl)
    I, J = M.shape
    lo, hi = (int(x) for x in args.nsig_range.split(","))
    h0 = synth.hyper_tuple()
    tasks = []
    for n in range(lo, hi + 1):
        for c in range(args.chains):
            h = list(h0); h[3] = h0[3] / np.sqrt(n)                     # be/sqrt(n), R/signeR.R:163
            tasks.append(dict(K=n, hyper=tuple(h), burn=args.nsig_burn, eval=args.nsig_eval, seed=1000 * n + c))

    def runner(ts):                                                      # initial states only for this rank's share
        for t in ts:
            t["state"] = synth.init_state(I, J, t["K"], hyper=dict(be=t["hyper"][3]), seed=t["seed"])
        return batch_bics(M, W, ts, devices=[local]) if ts else []
    runner([dict(tasks[0], burn=5, eval=5)])                             # warm-up (context, module load, cohort upload)
    dist.barrier()
    t0 = time.perf_counter()
    res = dist.run_tasks_distributed(tasks, runner)
    dist.barrier()
    dt = dist.max_over_ranks(time.perf_counter() - t0)
    if rank == 0:
        med = {}
        for tk, (ll, bic) in zip(tasks, res):
            med.setdefault(tk["K"], []).append(float(np.median(bic)))
        med = {k: float(np.mean(v)) for k, v in med.items()}
        best = max(sorted(med), key=lambda k: med[k])
        print(json.dumps(dict(metric="nsig_sweep_wall_s", value=dt, unit="s", n_gpus=world, higher_is_better=False, scaling="strong",
                              config=dict(workload=wl, I=I, J=J, nsig=[lo, hi], chains=args.chains, burn=args.nsig_burn, eval=args.nsig_eval, tasks=len(tasks)),
                              bic_optimal_nsig=best, median_bic=med, sweeps_per_s=sum(t["burn"] + t["eval"] for t in tasks) / dt)), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def run_sharded(args):
    """Secondary metric: ONE chain with its genome columns sharded over N GPUs (BASELINE.json config 4: 1536 x 20,000,
    K = 30), driven by one process through signer_gibbs_run_sharded (NCCL all-reduce of the I x K statistics each sweep).
    Times the whole reference-shaped call (upload, burn-in sweeps, final state back), so the value is an end-to-end one."""
    import signer_b200 as sg
    from signer_b200 import synth
    wl = args.workload
    K = WORKLOADS[wl]["K"]
    M, W, _, _ = synth.named(wl)
    I, J = M.shape
    st = synth.init_state(I, J, K, seed=4)
    h = synth.hyper_tuple()
    devs = list(range(args.gpus))
    a = (st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], *h)
    from signer_b200 import _lib
    sg.GibbsSamplerCpp(M, W, None, *a, 2, 0, False, False, False, False, False, seed=1, shard_devices=devs)     # warm-up
    t0 = time.perf_counter()
    r = sg.GibbsSamplerCpp(M, W, None, *a, args.steps, 0, False, False, False, False, False, seed=1, shard_devices=devs)
    dt = time.perf_counter() - t0
    loop_ms = float(_lib.lib().signer_last_loop_ms())          # the library's own clock around the sweep loop
    print(json.dumps(dict(metric="sharded_gibbs_sweeps_per_s", value=args.steps / (loop_ms * 1e-3), unit="sweeps/s", n_gpus=args.gpus,
                          steps=args.steps, ms_per_step=loop_ms / args.steps, higher_is_better=True, scaling="strong",
                          config=dict(workload=wl, I=I, J=J, K=K, shards=args.gpus, collective="ncclAllReduce of Zin (int32) and W E' sums (f64), 2*I*K values per sweep"),
                          call_seconds=dt, upload_and_setup_seconds=dt - loop_ms * 1e-3, zin_total=int(r.Zin.sum()), m_total=int(M.sum()))), flush=True)


def leg_em_loop(sg, M, W, st, hyper, Z0, local, calls=10, burn=200, ev=100):
    """The EM loop as R drives it through the drop-in boundary (R/cpp_eBayesNMF.R:104-152): every iteration is one
    stateless GibbsSamplerCpp(..., burn=200, eval=EM_eval_it=100, keep_par=TRUE) with host arrays, and the next one
    starts from the last samples of Zs, Ps, Es, Aps, Bps, Aes, Bes it returned (:108-139)."""
    P, E, Ap, Bp, Ae, Be = (st[k] for k in ("P", "E", "Ap", "Bp", "Ae", "Be"))
    Z = Z0
    sg.GibbsSamplerCpp(M, W, Z, P, E, Ap, Bp, Ae, Be, *hyper, 3, 3, False, False, False, False, True, seed=40, device=local)
    t0 = time.perf_counter()
    for it in range(calls):
        r = sg.GibbsSamplerCpp(M, W, Z, P, E, Ap, Bp, Ae, Be, *hyper, burn, ev, False, False, False, False, True, seed=41 + it,
                               device=local)
        Z = r[0][:, :, :, 0]
        P, E, Ap, Bp, Ae, Be = r[2][:, :, -1], r[3][:, :, -1], r[4][:, :, -1], r[5][:, :, -1], r[6][:, :, -1], r[7][:, :, -1]
    dt = time.perf_counter() - t0
    I, J = M.shape
    K = P.shape[1]
    per_call_h2d = 8.0 * (2 * I * J + I * J * K + 3 * (I * K + K * J))
    per_call_d2h = 8.0 * (I * J * K + ev * 3 * (I * K + K * J))
    return dict(calls=calls, burn=burn, eval=ev, keep_par=True, seconds=dt, seconds_per_call=dt / calls,
                sweeps_per_s=calls * (burn + ev) / dt, h2d_bytes_per_call=per_call_h2d, d2h_bytes_per_call=per_call_d2h,
                call="GibbsSamplerCpp(..., burn=200, eval=100, keep_par=TRUE) x %d, state fed back through host arrays "
                     "(R/cpp_eBayesNMF.R:104-152)" % calls)


def leg_nsig(M, W, local, lo=2, hi=20, chains=4, burn=1000, ev=1000):
    """nsig = lo..hi x `chains` chains, `burn` + `ev` sweeps each (R/signeR.R:66, be/sqrt(n) :163), dealt longest-first
    over the ranks (one process per GPU, no data-path collective; only BIC vectors travel).  Returns the wall time
    (max over ranks) and the BIC-optimal nsig."""
    from signer_b200 import dist as sgdist, synth
    from signer_b200.signer import batch_bics
    I, J = M.shape
    h0 = synth.hyper_tuple()
    tasks = []
    for n in range(lo, hi + 1):
        for c in range(chains):
            h = list(h0); h[3] = h0[3] / np.sqrt(n)
            tasks.append(dict(K=n, hyper=tuple(h), burn=burn, eval=ev, seed=1000 * n + c))

    def runner(ts):                                                      # initial states only for this rank's share
        for t in ts:
            t["state"] = synth.init_state(I, J, t["K"], hyper=dict(be=t["hyper"][3]), seed=t["seed"])
        return batch_bics(M, W, ts, devices=[local]) if ts else []
    sgdist.barrier()
    t0 = time.perf_counter()
    res = sgdist.run_tasks_distributed(tasks, runner)
    sgdist.barrier()
    dt = sgdist.max_over_ranks(time.perf_counter() - t0)
    med = {}
    for tk, (ll, bic) in zip(tasks, res):
        med.setdefault(tk["K"], []).append(float(np.median(bic)))
    med = {k: float(np.mean(v)) for k, v in med.items()}
    best = max(sorted(med), key=lambda k: med[k])
    return dict(nsig_wall_s=dt, bic_optimal_nsig=best, nsig=[lo, hi], chains=chains, burn=burn, eval=ev, tasks=len(tasks),
                sweeps_per_s=sum(t["burn"] + t["eval"] for t in tasks) / dt, scaling="strong",
                median_bic={str(k): v for k, v in sorted(med.items())})


def leg_sharded(n_dev, steps=20):
    """Rank 0 only.  (a) BASELINE.json config 4 (1536 x 20,000, K = 30): one chain, genome columns sharded over n_dev
    GPUs through signer_gibbs_run_sharded; ms per sweep from the library's own clock around the sweep loop.
    (b) a small sharded chain over the same devices (real NCCL when n_dev > 1; two emulated shards on the one GPU
    otherwise) compared bit for bit with the oracle run with the same number of shards -- the oracle is the checker
    here, never the thing measured."""
    import signer_b200 as sg
    from signer_b200 import _lib, synth
    wl = "cfg4_1536x20000_K30"
    K = WORKLOADS[wl]["K"]
    M, W, _, _ = synth.named(wl)
    I, J = M.shape
    st = synth.init_state(I, J, K, seed=4)
    h = synth.hyper_tuple()
    devs = list(range(n_dev))
    a = (st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], *h)
    sg.GibbsSamplerCpp(M, W, None, *a, 3, 0, False, False, False, False, False, seed=1, shard_devices=devs)     # warm-up
    t0 = time.perf_counter()
    r = sg.GibbsSamplerCpp(M, W, None, *a, steps, 0, False, False, False, False, False, seed=1, shard_devices=devs)
    dt = time.perf_counter() - t0
    loop_ms = float(_lib.lib().signer_last_loop_ms())
    out = dict(workload=wl, I=I, J=J, K=K, shards=n_dev, steps=steps, sharded_ms_per_sweep=loop_ms / steps,
               sharded_sweeps_per_s=steps / (loop_ms * 1e-3), call_seconds=dt, scaling="strong",
               checksum_ok=bool(int(r.Zin.sum()) == int(np.trunc(M).sum())),
               exchange="per sweep: Zin (int32, I*K) all-reduce + W E' (f64, I*K per shard) all-gather over NVLink" if n_dev > 1 else "none (one shard)")
    del M, W, r
    # (b) parity on a small cohort, all outputs
    from oracle import oracle as O
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import compare_all
    pdevs = devs if n_dev > 1 else [0, 0]
    I2, J2, K2 = 24, 128 * 9 + 17, 5
    M2, W2, _, _ = synth.cohort(I2, J2, K2, burden=25.0 * I2, seed=J2)
    M2[1, 3] = 4000.0
    st2 = synth.init_state(I2, J2, K2, seed=5)
    got = sg.GibbsSamplerCpp(M2, W2, None, st2["P"], st2["E"], st2["Ap"], st2["Bp"], st2["Ae"], st2["Be"], *h, 10, 5,
                             False, False, False, False, True, seed=9, shard_devices=pdevs)
    hd = dict(zip(("ap", "bp", "ae", "be", "lp", "le", "var_ap", "var_ae"), h))
    want = O.gibbs_run(M2, W2, None, st2["P"], st2["E"], st2["Ap"], st2["Bp"], st2["Ae"], st2["Be"], hyper=hd, burn=10, eval=5,
                       keep_par=True, seed=9, nshards=len(pdevs))
    try:
        compare_all(got, want)
        out["sharded_parity_ok"] = True
    except AssertionError as e:
        out["sharded_parity_ok"] = False
        out["sharded_parity_error"] = str(e)[:300]
    out["sharded_parity_case"] = "%dx%d K=%d, 10+5 sweeps, %d shards %s, every output bit for bit against the oracle" % (
        I2, J2, K2, len(pdevs), "over NCCL" if n_dev > 1 else "emulated on one GPU")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: cfg3_96x10000_K20 (sweeps), cfg2_96x560 (nsig), cfg4_1536x20000_K30 (sharded)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the em_loop, nsig and sharded legs (kernel work: faster turn-around)")
    ap.add_argument("--mode", default="sweeps", choices=["sweeps", "nsig", "sharded"],
                    help="nsig: wall time of a full nsig model-selection sweep (secondary metric of BASELINE.json)")
    ap.add_argument("--nsig-range", default="2,10")
    ap.add_argument("--chains", type=int, default=1)
    ap.add_argument("--nsig-burn", type=int, default=1000)
    ap.add_argument("--nsig-eval", type=int, default=1000)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.workload is None:
        args.workload = {"sweeps": "cfg3_96x10000_K20", "nsig": "cfg2_96x560", "sharded": "cfg4_1536x20000_K30"}[args.mode]

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.mode == "nsig":
        run_nsig(args)
        return
    if args.mode == "sharded":
        run_sharded(args)
        return

    import torch
    import torch.distributed as dist
    import signer_b200 as sg
    from signer_b200 import dist as sgdist, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    sgdist.init("nccl")
    dist_on = world > 1
    barrier = sgdist.barrier
    max_over_ranks = sgdist.max_over_ranks

    wl = WORKLOADS[args.workload]
    K = wl["K"]
    M, W, _, _ = synth.named(args.workload)
    I, J = M.shape
    st = synth.init_state(I, J, K, seed=1000 * K + rank)
    hyper = synth.hyper_tuple()
    chain = sg.Chain(M, W, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], hyper, seed=1000 * K + rank, device=local)
    # a dedicated (non-default) torch stream: the chain launches on it, torch events time it
    stream = torch.cuda.Stream(device=local)
    assert stream.cuda_stream != 0
    chain.set_stream(stream.cuda_stream)

    # ---------------- value: burn-in sweeps, device resident
    chain.run(burn=max(args.warmup, MIN_WARMUP_SWEEPS), fetch=False)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = chain.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    chain.run(burn=args.steps, fetch=False)
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = chain.launches - l0
    clk = clocks.stop() if rank == 0 else None
    value = world * args.steps / (ms * 1e-3)

    # ---------------- evaluation sweeps (samples + log-likelihood kept on the device)
    n_eval = min(args.steps, 500)
    chain.run(eval=n_eval, keep_par=False, fetch=False)      # warm-up: also sizes the device rings
    barrier()
    e0.record(stream)
    chain.run(eval=n_eval, keep_par=False, fetch=False)
    e1.record(stream)
    barrier()
    ms_eval = max_over_ranks(e0.elapsed_time(e1))
    eval_value = world * n_eval / (ms_eval * 1e-3)

    # ---------------- per-kernel times: second pass of burn-in sweeps with events around every launch
    n_prof = min(args.steps, 500)
    chain.profile(True)
    chain.run(burn=n_prof, fetch=False)
    prof = chain.profile(False)
    # ... and of evaluation sweeps (ring stores + log-likelihood), same way
    chain.profile(True)
    chain.run(eval=min(n_prof, 200), keep_par=False, fetch=False)
    prof_eval = chain.profile(False)
    kz = prof["z_alloc_fused"]
    z_ms = kz["ms"] / max(1, kz["launches"])
    tot_prof_ms = sum(v["ms"] for v in prof.values())
    z_bytes = 4.0 * I * J + 12.0 * K * (I + J)                  # M int32 in; P,E in; Zin,Znj out
    peak, peak_src = peaks()
    achieved = z_bytes / (z_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "z_alloc_fused_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = dict(bound="hbm", kernel="z_alloc_fused", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                    traffic=traffic, peak_source=peak_src, algorithmic_bytes_per_launch=z_bytes, ms_per_launch=z_ms,
                    share_of_sweep=kz["ms"] / tot_prof_ms if tot_prof_ms else None,
                    timing="second pass of %d burn-in sweeps with CUDA events around every launch" % n_prof,
                    note="issue/RNG-bound, working set L2-resident (SURVEY.md 8d): HBM fraction reported as defined; "
                         "see draws_per_s for the rate that binds")
    # actual step-1 draws per sweep and the measured ceiling of the uniform-draw rate
    chain.draw_counts(True)
    n_cnt = 50
    chain.run(burn=n_cnt, fetch=False)
    n_cat, n_bin = chain.draw_counts(False)
    from signer_b200 import _lib as _sl
    import ctypes as _C
    ceil = _C.c_double(0.0)
    _sl.check(_sl.lib().signer_draw_ceiling(local, _C.byref(ceil)))
    b_sweep = 20.0 * I * J + 104.0 * K * (I + J)
    per_rank_sps = args.steps / (ms * 1e-3)
    kernels_ms = {k: (v["ms"] / v["launches"] if v["launches"] else None) for k, v in prof.items()}
    chain.set_stream(None)

    # ---------------- e2e: the reference-shaped stateless call with host arrays
    e2e = None
    if not args.no_e2e:
        zrun = sg.GibbsSamplerCpp(M, W, None, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], *hyper, 0, 1,
                                  False, False, False, False, True, seed=5, device=local)
        Z0 = np.asfortranarray(zrun[0][:, :, :, 0])
        del zrun
        nb = max(1, args.steps // 2)
        ne = max(1, args.steps - nb)
        wb = max(3, args.warmup // 2)
        sg.GibbsSamplerCpp(M, W, Z0, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], *hyper, wb, max(3, args.warmup - wb),
                           False, False, False, False, False, seed=6, device=local)      # the W warm-up sweeps through the same call
        # three full calls, the median reported (all three kept in `seconds_all`): the call returns ~3 GB of freshly
        # allocated host pages, whose first-touch cost varies from call to call on the same box
        dts = []
        for rep in range(3):
            barrier()
            t0 = time.perf_counter()
            res = sg.GibbsSamplerCpp(M, W, Z0, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], *hyper, nb, ne,
                                     False, False, False, False, False, seed=7, device=local)
            torch.cuda.synchronize()
            dts.append(max_over_ranks(time.perf_counter() - t0))
            if rep < 2:
                del res
        dt = float(np.median(dts))
        h2d = 8.0 * (2 * I * J + I * J * K + 3 * (I * K + K * J)) + 4.0 * I * J * 0
        d2h = 8.0 * ne * 2 * (I * K + K * J) + 16.0 * ne + 8.0 * 3 * (I * K + K * J) + 4.0 * (I * K + K * J)
        e2e = dict(value=world * (nb + ne) / dt, unit=UNIT, h2d_bytes_per_step=h2d / (nb + ne), d2h_bytes_per_step=d2h / (nb + ne),
                   call="GibbsSamplerCpp(M,W,Z,P,E,Ap,Bp,Ae,Be,...,burn=%d,eval=%d,keep_par=FALSE) with host arrays; "
                        "pageable host memory as R hands it over; median of three calls" % (nb, ne), seconds=dt, seconds_all=dts,
                   median_bic=float(np.median(res.bic)))
        del res

    # ---------------- extras: the EM-shaped loop, the nsig sweep and the column-sharded chain (docstring)
    em = nsig = sharded = None
    if not args.no_extras and args.workload == "cfg3_96x10000_K20":
        if rank == 0 and not args.no_e2e:
            em = leg_em_loop(sg, M, W, st, hyper, Z0, local)
        chain.close()
        sg.release_caches()
        nsig = leg_nsig(M, W, local)
        # rank 0 drives all N devices; the other ranks wait on the HOST (a gloo barrier): an NCCL barrier would park a
        # spinning kernel on the very GPUs the sharded chain needs
        cpu_group = dist.new_group(backend="gloo") if dist_on else None
        if rank == 0:
            try:
                sharded = leg_sharded(world)
            except Exception as e:                      # keep the headline line even if this leg fails
                sharded = dict(error=str(e)[:300])
            sg.release_caches()
        if dist_on:
            dist.barrier(group=cpu_group)

    # ---------------- CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, kind, sample = cpu_sweep_rate(M, W, K, seconds_target=15.0)
        cpu = dict(value=v, unit=UNIT, cores=1, kind=kind, sample=sample)

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                    data="synthetic",
                    config=config_dict(args.workload, I, J, K),
                    clocks=clk, e2e=e2e, gpu_launches=int(launches) * world,
                    eval_value=eval_value, eval_ms_per_step=ms_eval / n_eval,
                    roofline=roofline,
                    roofline_sweep=dict(algorithmic_bytes_per_sweep=b_sweep, achieved=b_sweep * per_rank_sps / 1e9, peak=peak,
                                        unit="GB/s", frac=b_sweep * per_rank_sps / 1e9 / peak),
                    kernel_ms=kernels_ms,
                    kernel_ms_eval={k: (v["ms"] / v["launches"] if v["launches"] else None) for k, v in prof_eval.items()},
                    draws_per_s=dict(step1_categorical=per_rank_sps * n_cat / n_cnt, step1_binomial=per_rank_sps * n_bin / n_cnt,
                                     gamma=per_rank_sps * 3 * K * (I + J), mh_uniform=per_rank_sps * K * (I + J),
                                     step1_draws_per_sweep=(n_cat + n_bin) / n_cnt,
                                     reference_binomials_per_sweep_upper_bound=I * J * (K - 1),
                                     ceiling_uniform_draws_per_s=ceil.value,
                                     step1_fraction_of_ceiling=(n_cat + n_bin) / n_cnt / (z_ms * 1e-3) / ceil.value,
                                     note="ceiling: Philox4x32-10 + 32-bit uniform + one inversion step per lane, no memory (signer_draw_ceiling); "
                                          "fraction = step-1 draws per second of z_alloc_fused time / ceiling"),
                    cpu_baseline=cpu, em_loop=em, nsig=nsig, sharded=sharded)
        if nsig:
            line["nsig_wall_s"] = nsig["nsig_wall_s"]; line["bic_optimal_nsig"] = nsig["bic_optimal_nsig"]
        if sharded and "sharded_ms_per_sweep" in sharded:
            line["sharded_ms_per_sweep"] = sharded["sharded_ms_per_sweep"]; line["sharded_parity_ok"] = sharded.get("sharded_parity_ok")
        print(json.dumps(line), flush=True)
    chain.close()
    if dist_on:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
