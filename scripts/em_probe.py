import sys, time, json, os
sys.path.insert(0, '/root/repo')
import numpy as np
import bench
import signer_b200 as sg
from signer_b200 import synth
M, W, _, _ = synth.named("cfg3_96x10000_K20")
I, J = M.shape; K = 20
st = synth.init_state(I, J, K, seed=1)
h = synth.hyper_tuple()
Z0 = np.zeros((I, J, K), order="F"); Z0[:, :, 0] = np.trunc(M)
big = len(sys.argv) > 1 and sys.argv[1] == "big"
if big:   # what the default bench does before the EM leg: three 1000+1000 calls with 3.2 GB of samples each
    for i in range(3):
        r = sg.GibbsSamplerCpp(M, W, Z0, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], *h, 1000, 1000, False, False, False, False, False, seed=7 + i)
    del r
for rep in range(3):
    d = bench.leg_em_loop(sg, M, W, st, h, Z0, 0)
    print("big" if big else "small", os.environ.get("SIGNER_RING_BYTES", "-"), round(d["seconds_per_call"] * 1e3, 1), "ms per call", flush=True)
