"""Config 1 (the vignette's 21 breast genomes): BIC-optimal nsig with the reference's default budgets
(testing_burn = testing_eval = 1000, R/signeR.R:66) for several seeds and both initialisations.  GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import signer_b200 as sg
from signer_b200 import synth
M, W, _ = synth.breast21()
for start in ("lee", "random"):
    for seed in (42, 1, 2, 3):
        sg.set_seed(seed)
        res = sg.signeR(M.T, samples="rows", Opport=W, nlim=(2, 8), try_all=True, start=start, testing_burn=1000, testing_eval=1000,
                        main_burn=100, main_eval=50, rng=np.random.default_rng(seed))
        med = [round(float(np.median(b)), 1) for b in res["Test_BICs"]]
        print("start=%-6s seed=%-3d nopt=%d  median BIC n=2..8: %s" % (start, seed, res["Nsign"], med), flush=True)
