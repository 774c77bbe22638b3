#!/usr/bin/env python
"""Experiment (GPU box): per-batch wall cycles of z_alloc_fused's last launch, from a library built with -DSG_Z_TIMING.
usage: batch_cycles.py <lib.so> [workload] [K]  -> gpurun_out/batch_cycles.npy + a summary"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import shutil
shutil.copy(sys.argv[1], os.path.join(ROOT, "signer_b200", "libsigner_gibbs.so"))
import signer_b200 as sg
from signer_b200 import synth, _lib
wl = sys.argv[2] if len(sys.argv) > 2 else "cfg3_96x10000_K20"
K = int(sys.argv[3]) if len(sys.argv) > 3 else 20
M, W, _, _ = synth.named(wl)
I, J = M.shape
st = synth.init_state(I, J, K, seed=1)
ch = sg.Chain(M, W, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], synth.hyper_tuple(), seed=1000 * K)
ch.run(burn=300, fetch=False); ch.sync()
L = _lib.lib()
nb = ((M >= 1).sum() + 31) // 32
n = int(min(nb, 1 << 16))
out = np.zeros(n, dtype=np.uint32)
L.signer_debug_batch_cycles.argtypes = [C.c_void_p, C.c_int]
rc = L.signer_debug_batch_cycles(out.ctypes.data_as(C.c_void_p), n)
assert rc == 0, rc
np.save(os.path.join(ROOT, "gpurun_out", "batch_cycles.npy"), out)
cnt = np.sort(M[M >= 1].astype(np.int64))[::-1]
print("batches", n, "total Mcycles", out.sum() / 1e6, "max", out.max(), "sum/3552", out.sum() / 3552)
for lo, hi in [(0, 10), (10, 100), (100, 300), (300, 1000), (1000, 2000), (2000, 3552), (3552, 5000), (5000, 7200), (7200, 12000), (12000, 20000), (20000, n)]:
    hi = min(hi, n)
    if lo >= hi: break
    print("batches %6d..%6d  n %6d..%6d  mean cycles %9.0f  max %9d" % (lo, hi, cnt[min(lo * 32, cnt.size - 1)], cnt[min(hi * 32 - 1, cnt.size - 1)], out[lo:hi].mean(), out[lo:hi].max()))

ph = np.zeros((2, 8), dtype=np.uint64)
L.signer_debug_z_phases.argtypes = [C.c_void_p]
assert L.signer_debug_z_phases(ph.ctypes.data_as(C.c_void_p)) == 0
names = ["cells+claim", "staging", "direct loop", "sort", "chain", "tail", "degenerate", "-"]
tot = ph.sum()
for c, cname in enumerate(["small batches", "large batches"]):
    print(cname, " ".join("%s %.1f%%" % (names[i], 100.0 * ph[c, i] / tot) for i in range(7)), " | class total %.1f%%" % (100.0 * ph[c].sum() / tot))
