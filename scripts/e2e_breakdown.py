"""Where does the reference-shaped call spend its time?  (run on the GPU box)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import signer_b200 as sg
from signer_b200 import synth
M, W, _, _ = synth.named("cfg3_96x10000_K20")
I, J = M.shape; K = 20
st = synth.init_state(I, J, K, seed=1); h = synth.hyper_tuple()
def T(f, n=1):
    t = time.perf_counter(); r = None
    for _ in range(n): r = f()
    return (time.perf_counter() - t) / n, r
t, z = T(lambda: sg.GibbsSamplerCpp(M, W, None, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], *h, 0, 1, False, False, False, False, True, seed=5))
Z0 = np.asfortranarray(z[0][:, :, :, 0]); print("first call (context init etc) %.3f s" % t)
args = (st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], *h)
for (Zin, burn, ev, keep, label) in [(None, 0, 1, False, "Z=None burn0 eval1"), (Z0, 0, 1, False, "Z cube burn0 eval1"),
                                     (None, 1000, 0, False, "Z=None burn1000 eval0"), (None, 0, 1000, False, "Z=None burn0 eval1000 keep=F"),
                                     (Z0, 1000, 1000, False, "Z cube burn1000 eval1000 keep=F"), (Z0, 1000, 1000, True, "Z cube burn1000 eval1000 keep=T")]:
    t, _ = T(lambda: sg.GibbsSamplerCpp(M, W, Zin, *args, burn, ev, False, False, False, False, keep, seed=7), 2)
    print("%-40s %.3f s  -> %.0f sweeps/s" % (label, t, (burn + ev) / t))
t, c = T(lambda: sg.Chain(M, W, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], h, seed=3)); print("Chain create (upload, cell list) %.3f s" % t)
t, _ = T(lambda: c.run(burn=1000, fetch=False) or c.sync()); print("1000 burn resident %.3f s" % t)
t, _ = T(lambda: c.run(eval=1000, fetch=False) or c.sync()); print("1000 eval resident (first: ring alloc) %.3f s" % t)
t, _ = T(lambda: c.run(eval=1000, fetch=False) or c.sync()); print("1000 eval resident %.3f s" % t)
t, r = T(lambda: c.run(eval=1000, fetch=True, want_Z=False)); print("1000 eval + D2H into fresh numpy arrays %.3f s" % t)
import ctypes
buf = np.zeros((K, J, 1000), order="F"); t, _ = T(lambda: buf.fill(1.0)); print("host touch of one 1.6 GB array %.3f s" % t)
