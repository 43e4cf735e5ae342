"""Posterior parity at SURVEY.md section 8's stated tolerances, on BASELINE.json's configurations:

  every marginal: two-sample KS  D < 1.95 / sqrt(n_eff / 2)  (alpha ~ 0.001), median and 2.5 / 97.5 % quantiles
  within 3 Monte-Carlo standard errors (tests/stat_helpers.py), for >= 99 % of the marginals.

(a) configs[0]: one event x 10 stations with the MCMC block of sample/hypo_tremor.in UNCHANGED (solve_* = T, the
    sample's step sizes, n_chains 5, n_cool 1, temp_high 200) and n_procs = 4 -- the float32 blocked-Gibbs kernel
    against the oracle's mode A (the reference's own schedule and mod_random), all 25 marginals
    (x, y, z, vs, qs, 10 t_corr, 10 a_corr).
(b) configs[1], configs[2] and a configs[3]-sized catalogue (100 000 events x 50 stations on one GPU) at FULL size on
    the GPU (float32 lane kernel) against the float64 oracle on a subset of 100 of the same events (events are independent when solve_* = F and Philox ids are global, so the oracle runs
    exactly those events' chains).
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import hypotremormcmc_b200 as H
from oracle.pyoracle import Oracle
from stat_helpers import compare_marginal

pytestmark = pytest.mark.gpu
NOSOLVE = dict(solve_vs=0, solve_t_corr=0, solve_qs=0, solve_a_corr=0)


def by_rank(samples, key, col=None):
    out = []
    for s in samples:
        v = s[key]
        out.append(np.asarray(v if col is None else v[:, col], dtype=np.float64))
    return out


def test_config0_sample_settings_float32_gibbs_vs_reference_schedule():
    E, S, R, K = 1, 10, 4, 5
    syn = H.Synthetic(E, S, 20231001)
    base = dict(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1)          # everything else: the sample file's values
    # reference schedule: ONE scalar per iteration (2.5 % each shared type, 90 % a hypocentre coordinate)
    n_a, burn_a, int_a = 12000000, 2000000, 200
    cfgA = H.default_config(mode=H.MODE_REPLAY, precision=64, n_iter=n_a, n_burn=burn_a, n_interval=int_a, **base)
    o = Oracle(cfgA, syn)
    o.init_chains()
    o.run_threaded(1, n_a)
    sa = [o.fetch_samples(r) for r in range(R)]
    # blocked Gibbs: every iteration moves the hypocentre AND one shared parameter of every chain
    n_c, burn_c, int_c = 1500000, 200000, 20
    cfgC = H.default_config(mode=H.MODE_BLOCKED_GIBBS, precision=32, n_iter=n_c, n_burn=burn_c, n_interval=int_c,
                            max_samples=10001, **base)
    parts = [[] for _ in range(R)]
    with H.HypoTremorB200(cfgC) as g:
        g.load(syn)
        g.init_chains()
        it0 = 1
        while it0 <= n_c:
            it1 = min(n_c, it0 + 10000 * int_c - 1)
            g.run(it0, it1)
            for r in range(R):
                parts[r].append(g.fetch_samples(r))
                g.fetch_likelihood(r)
            it0 = it1 + 1
    sc = [{k: np.concatenate([p[k] for p in parts[r]]) for k in ("vs", "qs", "hypo", "t_corr", "a_corr")} for r in range(R)]
    assert sum(len(s["vs"]) for s in sc) == R * ((n_c - burn_c) // int_c)
    marg = [("x", "hypo", 0), ("y", "hypo", 1), ("z", "hypo", 2), ("vs", "vs", None), ("qs", "qs", None)]
    marg += [("t_corr%d" % j, "t_corr", j) for j in range(S)] + [("a_corr%d" % j, "a_corr", j) for j in range(S)]
    bad = []
    for name, key, col in marg:
        # cold chains hop between ranks (temperatures are exchanged), so a rank's file is a union of chain segments:
        # per-rank series are still the right unit for the autocorrelation estimate
        r = compare_marginal(by_rank(sa, key, col), by_rank(sc, key, col))
        if not r["ok"]:
            bad.append((name, r))
    assert len(bad) <= len(marg) // 100, bad          # >= 99 % of 25 marginals = all of them


def subset(syn, lo, n):
    sub = syn.shard(0, 1)
    for name in ("true_x", "true_y", "true_z", "x_mu", "y_mu", "t_obs", "t_stdv", "a_obs", "a_stdv"):
        setattr(sub, name, np.ascontiguousarray(getattr(syn, name)[lo:lo + n]))
    sub.n_events = n
    return sub


@pytest.mark.parametrize("E,S,n_it,burn,interval", [(1000, 20, 12000, 2000, 5), (10000, 50, 6000, 1000, 5),
                                                    (100000, 50, 6000, 1000, 5)],
                         ids=["configs1", "configs2", "configs3_size_on_one_gpu"])
def test_full_size_float32_vs_float64_oracle_on_100_events(E, S, n_it, burn, interval):
    R, K, n_sub, n_blocks, chunk = 4, 16, 100, 10, 50
    syn = H.Synthetic(E, S, 20231001 + (2 if E == 1000 else 3))
    cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=n_it, n_burn=burn,
                           n_interval=interval, mode=H.MODE_FACTORISED, precision=32, seed=77,
                           max_samples=chunk + 1, **NOSOLVE)
    picks = [(E // n_blocks) * b + 7 for b in range(n_blocks)]      # ten blocks of ten events spread over the catalogue
    nb = n_sub // n_blocks
    cols = np.concatenate([np.arange(3 * lo, 3 * (lo + nb)) for lo in picks])

    def run_block(lo):
        o = Oracle(H.copy_config(cfg, precision=64, n_events=nb, max_samples=0), subset(syn, lo, nb), event_offset=lo)
        o.init_chains()
        o.run(1, n_it, trace=False)
        return [o.fetch_samples(r)["hypo"] for r in range(R)]

    with ThreadPoolExecutor(max_workers=n_blocks) as pool:        # the oracle call releases the GIL
        fut = [pool.submit(run_block, lo) for lo in picks]
        parts = [[] for _ in range(R)]
        with H.HypoTremorB200(cfg) as g:
            g.load(syn)
            g.init_chains()
            it0 = 1
            while it0 <= n_it:                                     # drained in chunks like the driver does; only the
                it1 = min(n_it, it0 + chunk * interval - 1)        # subset's columns are kept
                g.run(it0, it1)
                for r in range(R):
                    parts[r].append(g.fetch_samples(r)["hypo"][:, cols].copy())
                    g.fetch_likelihood(r)
                it0 = it1 + 1
        sg = [np.concatenate(p) for p in parts]
        so = [f.result() for f in fut]
    assert sg[0].shape == ((n_it - burn) // interval, 3 * n_sub)
    bad, n_marg = [], 0
    for b, lo in enumerate(picks):
        for e in range(nb):
            for c in range(3):
                n_marg += 1
                r = compare_marginal([so[b][rk][:, 3 * e + c] for rk in range(R)],
                                     [sg[rk][:, 3 * (b * nb + e) + c] for rk in range(R)])
                if not r["ok"]:
                    bad.append((lo + e, c, r["D"], r["D_max"], r["n_eff"]))
    assert n_marg == 300 and len(bad) <= 3, bad           # >= 99 % of the marginals


@pytest.mark.parametrize("mode", ["factorised", "blocked_gibbs"])
def test_device_posterior_quantiles_equal_the_sorted_samples(mode):
    """(f)-1: the device-side summary (radix selection on the marginal-major store, csrc/htm_summary.cu) must pick
    exactly the elements `hypo_tremor_statistics` picks from the sorted samples (src/cls_statistics.f90:230-232):
    compared bit for bit with hypotremormcmc_b200.io.quantile_summary on the records the same run handed out."""
    from hypotremormcmc_b200 import io as hio
    E, S, R, K = 37, 11, 3, 4
    syn = H.Synthetic(E, S, 5)
    n_it, burn, interval = 6000, 1500, 7
    kw = dict(NOSOLVE, mode=H.MODE_FACTORISED) if mode == "factorised" else dict(mode=H.MODE_BLOCKED_GIBBS)
    cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=2, n_iter=n_it, n_burn=burn,
                           n_interval=interval, precision=32, max_samples=120, summary=1, **kw)
    parts = []
    with H.HypoTremorB200(cfg) as g:
        g.load(syn)
        g.init_chains()
        it0 = 1
        while it0 <= n_it:
            it1 = min(n_it, it0 + 100 * interval - 1)
            g.run(it0, it1)
            for r in range(R):
                parts.append(g.fetch_samples(r))
                g.fetch_likelihood(r)
            it0 = it1 + 1
        q = g.posterior_quantiles()
    all_ = {k: np.concatenate([p[k] for p in parts]) for k in ("vs", "qs", "hypo", "t_corr", "a_corr")}
    n_mod = (n_it - burn) * R * cfg.n_cool // interval                      # src/cls_statistics.f90:65
    assert q["n"] == len(all_["vs"]) and abs(q["n"] - n_mod) <= R * cfg.n_cool
    for m in range(3 * E):
        assert tuple(q["hypo"][m]) == tuple(hio.quantile_summary(all_["hypo"][:, m])), m
    assert tuple(q["vs"]) == tuple(hio.quantile_summary(all_["vs"])) and tuple(q["qs"]) == tuple(hio.quantile_summary(all_["qs"]))
    for j in range(S):
        assert tuple(q["t_corr"][j]) == tuple(hio.quantile_summary(all_["t_corr"][:, j])), j
        assert tuple(q["a_corr"][j]) == tuple(hio.quantile_summary(all_["a_corr"][:, j])), j
    assert np.all(q["hypo"][:, 1] <= q["hypo"][:, 0]) and np.all(q["hypo"][:, 0] <= q["hypo"][:, 2])
