"""SignExp: the pair of sample tensors the sampler returns, with the reference's normalisation and
summaries (R/Definition_SignExp_object.R:2-61 class and constructor, :105-150 Normalize,
:247-303 Median_sign / Median_exp).  Only the part of the class on the sampler path is mirrored."""
import numpy as np


def _fivenum_whiskers(x, coef=1.5):
    """Tukey five-number summary with boxplot whiskers along the last axis, as R's
    boxplot.stats (hinges from fivenum, whiskers at the most extreme points within
    coef * IQR of the hinges).  x: (..., r) -> (..., 5)."""
    xs = np.sort(x, axis=-1)
    n = xs.shape[-1]
    n4 = np.floor((n + 3) / 2.0) / 2.0
    d = np.array([1.0, n4, (n + 1) / 2.0, n + 1 - n4, float(n)]) - 1.0
    lo, hi = np.floor(d).astype(int), np.ceil(d).astype(int)
    five = 0.5 * (xs[..., lo] + xs[..., hi])
    iqr = five[..., 3] - five[..., 1]
    lo_f = (five[..., 1] - coef * iqr)[..., None]
    hi_f = (five[..., 3] + coef * iqr)[..., None]
    inside = (xs >= lo_f) & (xs <= hi_f)
    big = np.where(inside, xs, np.inf).min(axis=-1)
    small = np.where(inside, xs, -np.inf).max(axis=-1)
    out = five.copy()
    out[..., 0] = np.where(np.isfinite(big), big, five[..., 0])
    out[..., 4] = np.where(np.isfinite(small), small, five[..., 4])
    return out


class SignExp:
    """Sign [i,n,r], Exp [n,j,r] (Fortran order like R arrays)."""

    def __init__(self, Ps, Es, samplenames=None, mutnames=None, signames=None):
        Ps = np.asfortranarray(np.asarray(Ps, dtype=np.float64))
        Es = np.asfortranarray(np.asarray(Es, dtype=np.float64))
        if Ps.ndim != 3 or Es.ndim != 3:
            raise ValueError("Signatures and exposures must be arrays.")
        if not (Ps.shape[2] == Es.shape[2] and Ps.shape[1] == Es.shape[0]):
            raise ValueError("Signatures and exposures are not compatible")
        i, n, r = Ps.shape
        j = Es.shape[1]
        self.sigSums = Ps.sum(axis=0)                                   # [n, r]  (:35)
        self.samples = list(samplenames) if samplenames is not None else ["sample_%d" % (g + 1) for g in range(j)]
        if mutnames is None:
            bases = "ACGT"
            mutnames = []
            for sub, refb, alts in (("C", "C", "AGT"), ("T", "T", "ACG")):
                for alt in alts:
                    for b5 in bases:
                        for b3 in bases:
                            mutnames.append("%s%s%s>%s" % (b5, refb, b3, alt))
            mutnames = mutnames[:i] if i <= 96 else ["m%d" % (s + 1) for s in range(i)]
        self.mutations = list(mutnames)
        self.signames = list(signames) if signames is not None else ["S%d" % (m + 1) for m in range(n)]
        self.Sign, self.Exp = Ps.copy(order="F"), Es.copy(order="F")
        self.normalized = False
        self.Psummary = self.Esummary = None
        self.normalize()

    def normalize(self):
        """R/Definition_SignExp_object.R:105-150."""
        if self.normalized:
            return self
        vm = self.sigSums                                               # [n, r]
        self.Sign = np.asfortranarray(self.Sign / vm[None, :, :])
        self.Exp = np.asfortranarray(self.Exp * vm[:, None, :])
        self.normalized = True
        self.Psummary = np.quantile(self.Sign, [0.05, 0.25, 0.50, 0.75, 0.95, 1.0], axis=2).transpose(1, 2, 0)
        self.Esummary = _fivenum_whiskers(self.Exp)
        return self

    def median_sign(self):
        return np.asfortranarray(self.Psummary[:, :, 2])               # :259

    def median_exp(self):
        return np.asfortranarray(self.Esummary[:, :, 2])               # :297

    def average_exp(self):
        return np.asfortranarray(self.Exp.mean(axis=2))
