"""signer_b200 -- B200 (sm_100a) implementation of signeR's empirical-Bayes Poisson-NMF Gibbs
sampler behind the reference's own boundary.  The CUDA library (libsigner_gibbs.so, C ABI in
include/signer_gibbs.h) is the product; this package is the host-side mirror of the reference's
R interface for that one path.  No CPU fallback exists."""
from .gibbs import Chain, GibbsSamplerCpp, SampleList, release_caches, set_seed   # noqa: F401
from ._lib import SignerError                                     # noqa: F401
from .ebayes import eBayesNMF                                     # noqa: F401
from .estimate import EstimateParameters                          # noqa: F401
from .signer import Optimal_sigs, signeR                          # noqa: F401
from .signexp import SignExp                                      # noqa: F401

__all__ = ["GibbsSamplerCpp", "Chain", "SampleList", "SignerError", "set_seed", "release_caches", "eBayesNMF", "EstimateParameters",
           "Optimal_sigs", "signeR", "SignExp"]
