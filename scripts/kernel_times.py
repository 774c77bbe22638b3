#!/usr/bin/env python
"""Per-kernel device times of burn-in and evaluation sweeps on a named workload (GPU box).
usage: kernel_times.py <workload> [K ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import signer_b200 as sg
from signer_b200 import synth
wl = sys.argv[1]
M, W, _, _ = synth.named(wl)
I, J = M.shape
Ks = [int(x) for x in sys.argv[2:]] or [dict(cfg2_96x560=10, cfg3_96x10000_K20=20, cfg4_1536x20000_K30=30, cfg5_96x100000=40)[wl]]
for K in Ks:
    st = synth.init_state(I, J, K, seed=K)
    n = 200 if I * J < 5e6 else 40
    with sg.Chain(M, W, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], synth.hyper_tuple(), seed=K) as c:
        c.run(burn=n, fetch=False); c.sync()
        import time
        t0 = time.perf_counter(); c.run(burn=n, fetch=False); c.sync(); dt = (time.perf_counter() - t0) / n
        c.profile(True); c.run(burn=n, fetch=False); pb = c.profile(False)
        c.run(eval=min(n, 50), fetch=False); c.profile(True); c.run(eval=min(n, 50), fetch=False); pe = c.profile(False)
    f = lambda p: {k: round(v["ms"] / max(1, v["launches"]), 4) for k, v in p.items()}
    print("%s I=%d J=%d K=%d  sweep %.4f ms | burn-in kernels %s | eval kernels %s" % (wl, I, J, K, dt * 1e3, f(pb), f(pe)), flush=True)
