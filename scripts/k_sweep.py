#!/usr/bin/env python
"""GPU box: per-kernel times of burn-in and evaluation sweeps on the config-3 cohort for several K.  usage: k_sweep.py K [K ...]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import signer_b200 as sg
from signer_b200 import synth
M, W, _, _ = synth.named("cfg3_96x10000_K20")
I, J = M.shape
for K in [int(a) for a in sys.argv[1:]]:
    st = synth.init_state(I, J, K, seed=1)
    ch = sg.Chain(M, W, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], synth.hyper_tuple(), seed=1000 * K)
    ch.run(burn=200, fetch=False); ch.sync()
    t0 = time.perf_counter(); ch.run(burn=500, fetch=False); ch.sync(); tb = (time.perf_counter() - t0) / 500
    ch.run(eval=100, fetch=False); ch.sync()
    t0 = time.perf_counter(); ch.run(eval=500, fetch=False); ch.sync(); te = (time.perf_counter() - t0) / 500
    ch.profile(True); ch.run(burn=200, fetch=False); pb = ch.profile(False)
    ch.profile(True); ch.run(eval=200, fetch=False); pe = ch.profile(False)
    f = lambda p: " ".join("%s %.1f" % (k[:6], 1e3 * v["ms"] / max(1, v["launches"])) for k, v in p.items())
    print("K=%2d  burn %.1f us/sweep  eval %.1f us/sweep | burn kernels(us): %s | eval kernels(us): %s" % (K, 1e6 * tb, 1e6 * te, f(pb), f(pe)), flush=True)
