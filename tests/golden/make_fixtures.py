"""Build tests/golden/breast21.npz from the reference's vignette data (run in the dev container,
where /root/reference exists; the GPU box only sees the committed .npz).

Source files (data, not code): inst/extdata/21_breast_cancers.mutations.txt (header + 21 genomes x
96 contexts), inst/extdata/21_breast_cancers.opportunity.txt (21 x 96, no header),
inst/extdata/Cosmic_signatures_BRC.txt (96 x 6).  Stored in the orientation signeR() uses
internally after its transposition (R/signeR.R:78-81,102-107): contexts x genomes.
"""
import os
import numpy as np

REF = "/root/reference/inst/extdata"
here = os.path.dirname(os.path.abspath(__file__))

with open(os.path.join(REF, "21_breast_cancers.mutations.txt")) as f:
    header = f.readline().rstrip("\n").split("\t")
    contexts = [h for h in header if h]
    rows = [line.rstrip("\n").split("\t") for line in f if line.strip()]
samples = [r[0] for r in rows]
M = np.array([[float(x) for x in r[1:]] for r in rows]).T            # 96 x 21
W = np.loadtxt(os.path.join(REF, "21_breast_cancers.opportunity.txt")).T  # 96 x 21
with open(os.path.join(REF, "Cosmic_signatures_BRC.txt")) as f:
    signames = f.readline().split()
    lines = [l.split() for l in f if l.strip()]
cosmic = np.array([[float(x) for x in l[1:]] for l in lines])          # 96 x 6
assert M.shape == (96, 21) and W.shape == (96, 21) and cosmic.shape == (96, 6)
np.savez_compressed(os.path.join(here, "breast21.npz"), M=M, W=W, cosmic=cosmic,
                    contexts=np.array(contexts), samples=np.array(samples), signames=np.array(signames))
print("wrote breast21.npz", M.sum(), W.mean())
