"""EstimateParameters (R/EstimateParameters.R:1-13): the M-step of the empirical-Bayes EM over the
hyper-hyperparameters.  l = 1/mean(As); (a, b) = the gamma maximum-likelihood estimate on all Bs.
The reference maximises the gamma log-likelihood numerically (nloptr SBPLX from the old values);
the maximiser only depends on three sufficient statistics (count, sum B, sum log B), which the
device reduces (signer_chain_fetch_stats), and is found here by Newton's method on
log a - digamma(a) = log(mean B) - mean(log B),  b = a / mean B."""
import numpy as np
from scipy import special


def gamma_mle(n, sum_b, sum_log_b, a0=1.0):
    mean_b = sum_b / n
    s = np.log(mean_b) - sum_log_b / n
    if not np.isfinite(s) or s <= 0:
        return float(a0), float(a0 / mean_b)
    a = (3.0 - s + np.sqrt((s - 3.0) ** 2 + 24.0 * s)) / (12.0 * s)      # Minka's starting point
    for _ in range(100):
        f = np.log(a) - special.digamma(a) - s
        fp = 1.0 / a - special.polygamma(1, a)
        step = f / fp
        a_new = a - step
        if a_new <= 0:
            a_new = a / 2.0
        if abs(a_new - a) <= 1e-14 * a:
            a = a_new
            break
        a = a_new
    return float(a), float(a / mean_b)


def EstimateParameters(As, Bs, aold, bold):
    """Same signature and return order as the reference: list(a, b, l)."""
    As = np.asarray(As, dtype=np.float64).ravel()
    Bs = np.asarray(Bs, dtype=np.float64).ravel()
    l = 1.0 / As.mean()
    a, b = gamma_mle(Bs.size, Bs.sum(), np.log(Bs).sum(), a0=aold)
    return [a, b, l]


def estimate_from_stats(stats4, aold):
    """stats4 = (sum A, sum B, sum log B, count) as returned by Chain.stats()."""
    sa, sb, slb, n = stats4
    a, b = gamma_mle(n, sb, slb, a0=aold)
    return [a, b, n / sa]
