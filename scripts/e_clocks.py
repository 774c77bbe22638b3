#!/usr/bin/env python
"""Experiment (GPU box): phase clocks of e_family's last launch (library built with -DSG_Z_TIMING).  usage: e_clocks.py <lib.so>"""
import ctypes as C, os, sys, shutil
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
shutil.copy(sys.argv[1], os.path.join(ROOT, "signer_b200", "libsigner_gibbs.so"))
import signer_b200 as sg
from signer_b200 import synth, _lib
K = 20
M, W, _, _ = synth.named("cfg3_96x10000_K20")
I, J = M.shape
st = synth.init_state(I, J, K, seed=1)
ch = sg.Chain(M, W, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], synth.hyper_tuple(), seed=1000 * K)
ch.run(burn=300, fetch=False); ch.sync()
nb = ((J + 127) // 128) * ((K + 3) // 4)
out = np.zeros((nb, 5), dtype=np.int64)
L = _lib.lib()
L.signer_debug_e_clocks.argtypes = [C.c_void_p, C.c_int]
assert L.signer_debug_e_clocks(out.ctypes.data_as(C.c_void_p), nb) == 0
t0 = out[:, 0].min()
d = out - t0
print("blocks", nb, " kernel span (cycles): start min/max", d[:, 0].min(), d[:, 0].max(), " end min/max", d[:, 4].min(), d[:, 4].max())
for a, b, name in [(0, 1, "phase A  P'W chain (thread 0)"), (1, 2, "phase B  element updates (thread 0)"), (2, 3, "barrier wait"), (3, 4, "phase C  W E' rows")]:
    x = out[:, b] - out[:, a]
    print("%-40s mean %8.0f  p10 %8.0f  p50 %8.0f  p90 %8.0f  max %8.0f" % (name, x.mean(), *np.percentile(x, [10, 50, 90]), x.max()))

# p_family: slots 0 start, 1/3/5 start of op o, 2/4/6 after the partial sums when op o is step 2, 7 end
pc = np.zeros((15, 8), dtype=np.int64)
L.signer_debug_p_clocks.argtypes = [C.c_void_p, C.c_int]
assert L.signer_debug_p_clocks(pc.ctypes.data_as(C.c_void_p), 15) == 0
import signer_b200
print("p_family thread-0 clocks relative to start (cycles), one row per block; last launch's op order unknown here:")
for b in range(0, 15, 5):
    print(b, (pc[b] - pc[b, 0]).tolist())
