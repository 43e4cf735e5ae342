"""Posterior-marginal comparison at the tolerances SURVEY.md section 8 states (the reference states none):
two-sample KS with D < 1.95 / sqrt(n_eff / 2) (alpha ~ 0.001) and median / 2.5 % / 97.5 % quantiles within 3
Monte-Carlo standard errors, where n_eff is the smaller effective sample size of the two (autocorrelated) series."""
import numpy as np


def autocorr_time(x, c=5.0):
    """integrated autocorrelation time (Sokal's automatic window: smallest M with M >= c * tau(M))"""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    if n < 8 or np.all(x == x[0]):
        return float(max(n, 1))
    x = x - x.mean()
    f = np.fft.rfft(x, 2 * n)
    acf = np.fft.irfft(f * np.conj(f))[:n]
    acf = acf / acf[0]
    tau = 2.0 * np.cumsum(acf) - 1.0
    m = np.arange(n)
    ok = m >= c * tau
    return float(max(1.0, tau[np.argmax(ok)] if ok.any() else tau[-1]))


def n_eff(chains):
    """chains: list of 1-D series (one per independent chain / rank); effective size of the pooled sample"""
    return float(sum(len(c) / autocorr_time(c) for c in chains))


def ks_distance(a, b):
    a, b = np.sort(a), np.sort(b)
    allv = np.concatenate([a, b])
    ca = np.searchsorted(a, allv, side="right") / a.size
    cb = np.searchsorted(b, allv, side="right") / b.size
    return float(np.max(np.abs(ca - cb)))


def compare_marginal(chains_a, chains_b, n_se=3.0):
    """-> dict(ok, D, D_max, n_eff, quantile checks).  chains_*: lists of per-chain series of ONE marginal."""
    a, b = np.concatenate(chains_a), np.concatenate(chains_b)
    ne = min(n_eff(chains_a), n_eff(chains_b))
    D = ks_distance(a, b)
    D_max = 1.95 / np.sqrt(ne / 2.0)
    pooled = np.sort(np.concatenate([a, b]))
    res = dict(D=D, D_max=D_max, n_eff=ne, ok=D < D_max, quantiles={})
    for q in (0.025, 0.5, 0.975):
        # a quantile estimate from n_eff independent draws has rank error sqrt(q (1 - q) / n_eff); two estimates
        # differ by sqrt(2) of that; translate n_se of them into values through the pooled quantile function
        dq = n_se * np.sqrt(2.0 * q * (1.0 - q) / ne)
        lo, hi = np.quantile(pooled, max(0.0, q - dq)), np.quantile(pooled, min(1.0, q + dq))
        qa, qb = np.quantile(a, q), np.quantile(b, q)
        good = abs(qa - qb) <= (hi - lo) + 1e-12 * max(1.0, abs(qa))
        res["quantiles"][q] = (qa, qb, hi - lo, bool(good))
        res["ok"] = res["ok"] and bool(good)
    return res
