"""The reference-shaped stateless call at the driver's bench size (burn = eval = 10) and the EM-shaped call
(burn 200 + eval 100, keep_par), with the library's phase trace (SIGNER_TRACE=1).  Run on the GPU box."""
import os, sys, time
os.environ.setdefault("SIGNER_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import signer_b200 as sg
from signer_b200 import synth
M, W, _, _ = synth.named("cfg3_96x10000_K20")
I, J = M.shape; K = 20
st = synth.init_state(I, J, K, seed=1); h = synth.hyper_tuple()
args = (st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], *h)
z = sg.GibbsSamplerCpp(M, W, None, *args, 0, 1, False, False, False, False, True, seed=5)
Z0 = np.asfortranarray(z[0][:, :, :, 0])
for (Zin, burn, ev, keep, label) in [(Z0, 10, 10, False, "cube in, 10+10 keep=F"), (None, 10, 10, False, "Z=None 10+10 keep=F"),
                                     (Z0, 200, 100, True, "cube in, 200+100 keep=T (EM call)"), (Z0, 1000, 1000, False, "cube in, 1000+1000 keep=F")]:
    for rep in range(4):
        t = time.perf_counter()
        r = sg.GibbsSamplerCpp(M, W, Zin, *args, burn, ev, False, False, False, False, keep, seed=7)
        dt = time.perf_counter() - t
        print("%-36s rep %d: %.2f ms -> %.0f sweeps/s" % (label, rep, dt * 1e3, (burn + ev) / dt), flush=True)
        del r
