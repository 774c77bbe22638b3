#!/usr/bin/env python
"""Experiment (GPU box): phase clocks of e_family's last launch (library built with -DSG_Z_TIMING: thread 0 of every block
stamps clock64() at the start, around the tile wait, after pass 1 (P'W), after the element updates and after pass 2 (W E') of its LAST round, and at the end).
usage: e_clocks.py <lib.so> [K]"""
import ctypes as C, os, sys, shutil
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
shutil.copy(sys.argv[1], os.path.join(ROOT, "signer_b200", "libsigner_gibbs.so"))
import signer_b200 as sg
from signer_b200 import synth, _lib
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
M, W, _, _ = synth.named("cfg3_96x10000_K20")
I, J = M.shape
st = synth.init_state(I, J, K, seed=1)
ch = sg.Chain(M, W, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], synth.hyper_tuple(), seed=1000 * K)
ch.run(burn=300, fetch=False); ch.sync()
nb = 4096
out = np.zeros((nb, 8), dtype=np.int64)
L = _lib.lib()
L.signer_debug_e_clocks.argtypes = [C.c_void_p, C.c_int]
assert L.signer_debug_e_clocks(out.ctypes.data_as(C.c_void_p), nb) == 0
out = out[out[:, 4] > 0]
t0 = out[:, 0].min()
d = out - t0
print("K", K, "blocks", len(out), " kernel span (cycles): start min/max", d[:, 0].min(), d[:, 0].max(), " end min/max", d[:, 4].min(), d[:, 4].max())
for a, b, name in [(0, 5, "kernel entry .. barriers initialised"), (5, 6, "round set-up, tile requests, prefetches"), (6, 7, "waiting for the tile"), (7, 1, "pass 1 (P'W, 96-step fma chains)"),
                   (1, 2, "element updates (warp 0's two elements)"), (2, 3, "pass 2 (W E' of the segment)"), (3, 4, "rest")]:
    x = out[:, b] - out[:, a]
    print("%-44s mean %8.0f  p10 %8.0f  p50 %8.0f  p90 %8.0f  max %8.0f" % (name, x.mean(), *np.percentile(x, [10, 50, 90]), x.max()))
