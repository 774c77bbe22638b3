"""BASELINE.json configs 1, 2, 5 as single-chain data points (run on the GPU box)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import signer_b200 as sg
from signer_b200 import synth
from oracle import oracle as O
try:
    from oracle import ref_runner
    REF = ref_runner if ref_runner.available() else O
except Exception:
    REF = O
out = {}
# config 1: the vignette data, nsig = 5, the reference's default final-fit budget 10000 + 2000
M, W, _ = synth.breast21()
st = synth.init_state(96, 21, 5, seed=1); h = synth.hyper_tuple()
a = (st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], *h)
sg.GibbsSamplerCpp(M, W, None, *a, 10, 10, False, False, False, False, True, seed=1)
t = time.perf_counter(); r = sg.GibbsSamplerCpp(M, W, None, *a, 10000, 2000, False, False, False, False, True, seed=1); dt = time.perf_counter() - t
t = time.perf_counter(); REF.gibbs_run(M, W, None, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], burn=300, eval=60, seed=1); dr = (time.perf_counter() - t) * 12000 / 360
out["cfg1_vignette_96x21_K5_10000+2000"] = dict(gpu_call_s=dt, gpu_sweeps_per_s=12000 / dt, cpu_reference_s_extrapolated=dr, cpu_sweeps_per_s=12000 / dr, median_bic=float(np.median(r.bic)))
for name, K, nsw in (("cfg2_96x560", 10, 2000), ("cfg5_96x100000", 40, 150)):
    M, W, _, _ = synth.named(name)
    I, J = M.shape
    st = synth.init_state(I, J, K, seed=2)
    with sg.Chain(M, W, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], h, seed=2) as c:
        c.run(burn=20, fetch=False); c.sync()
        t = time.perf_counter(); c.run(burn=nsw, fetch=False); c.sync(); dt = time.perf_counter() - t
        c.profile(True); c.run(burn=50, fetch=False); pr = c.profile(False)
    b_sweep = 20.0 * I * J + 104.0 * K * (I + J)
    out[name + "_K%d" % K] = dict(ms_per_sweep=dt / nsw * 1e3, sweeps_per_s=nsw / dt, kernel_ms={k: v["ms"] / max(1, v["launches"]) for k, v in pr.items()},
                                  b_sweep_MB=b_sweep / 1e6, hbm_GBs=b_sweep / (dt / nsw) / 1e9)
print(json.dumps(out, indent=1))
