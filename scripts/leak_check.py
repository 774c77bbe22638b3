#!/usr/bin/env python
"""Run ON THE GPU BOX: repeat the stateless call, the chain handle and the batch entry a few hundred times and watch
device memory (cudaMemGetInfo) and the process's resident set.  Prints one line per checkpoint; the numbers must level
off after the first checkpoints (pool, pinned staging and cohort cache warm)."""
import os, sys, resource
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import signer_b200 as sg
from signer_b200 import synth


def rss_mb():
    with open("/proc/self/statm") as f:
        return int(f.read().split()[1]) * os.sysconf("SC_PAGE_SIZE") / 2**20


def report(tag):
    free, total = torch.cuda.mem_get_info(0)
    print("%-34s device used %8.1f MB   host RSS %8.1f MB" % (tag, (total - free) / 2**20, rss_mb()), flush=True)


M, W, _, _ = synth.named("cfg2_96x560")
I, J = M.shape
h = synth.hyper_tuple()
report("start")
for rep in range(5):
    for i in range(60):
        K = 3 + (i % 6)
        st = synth.init_state(I, J, K, seed=i)
        Z = np.zeros((I, J, K), order="F"); Z[:, :, 0] = np.trunc(M)
        r = sg.GibbsSamplerCpp(M, W, Z, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], *h, 20, 10,
                               False, False, False, False, bool(i & 1), seed=i)
    del r
    report("stateless calls x%d" % (60 * (rep + 1)))
for rep in range(3):
    for i in range(60):
        K = 3 + (i % 6)
        st = synth.init_state(I, J, K, seed=i)
        with sg.Chain(M, W, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], h, seed=i) as c:
            c.run(burn=10, eval=5, keep_par=True)
            c.stats()
    report("chain handles x%d" % (60 * (rep + 1)))
from signer_b200.signer import batch_bics
for rep in range(3):
    tasks = [dict(K=K, state=synth.init_state(I, J, K, seed=K), hyper=h, burn=20, eval=10, seed=100 * rep + K) for K in range(2, 8) for _ in range(2)]
    batch_bics(M, W, tasks)
    report("batch runs x%d" % (rep + 1))
# a different cohort every call: the cohort cache must stay bounded
for rep in range(3):
    for i in range(30):
        M2 = M.copy(); M2[0, 0] += i + 30 * rep + 1
        st = synth.init_state(I, J, 4, seed=i)
        sg.GibbsSamplerCpp(M2, W, None, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], *h, 5, 2,
                           False, False, False, False, False, seed=i)
    report("new cohort every call x%d" % (30 * (rep + 1)))
sg.release_caches()
report("after release_caches()")
