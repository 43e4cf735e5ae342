"""Statistical pins of the oracle: with no data the cold chains must
sample the priors; and the factorised schedule (mode B) must have the same cold-chain posterior
as the reference schedule (mode A) when the shared parameters are fixed."""
import numpy as np
from scipy import stats

import hypotremormcmc_b200 as H
from oracle.pyoracle import Oracle

NOSOLVE = dict(solve_vs=0, solve_t_corr=0, solve_qs=0, solve_a_corr=0)


def cold_samples(o, R):
    s = [o.fetch_samples(r) for r in range(R)]
    return {k: np.concatenate([x[k] for x in s]) for k in ("hypo", "vs", "qs", "t_corr", "a_corr")}


def ks_ok(sample, cdf, thin, alpha=1e-3):
    return stats.kstest(sample[::thin], cdf).pvalue > alpha


def test_prior_only_reference_schedule():
    # use_time = use_amp = F  =>  L == 0  =>  cold marginals are the priors
    syn = H.Synthetic(1, 6, 21)
    cfg = H.default_config(n_sta=6, n_events=1, n_procs=2, n_chains=3, n_cool=1, n_iter=400000, n_burn=2000,
                           n_interval=7, mode=H.MODE_REPLAY, precision=64, use_time=0, use_amp=0,
                           step_size_xy=45.0, step_size_z=12.0, step_size_vs=1.5, step_size_qs=150.0,
                           step_size_t_corr=0.8, step_size_a_corr=0.03)
    o = Oracle(cfg, syn)
    o.init_chains()
    o.run(1, cfg.n_iter, trace=False)
    s = cold_samples(o, 2)
    n = len(s["vs"])
    assert n > 50000
    thin = 40
    assert ks_ok(s["hypo"][:, 0], stats.norm(syn.x_mu[0], cfg.prior_width_xy).cdf, thin)
    assert ks_ok(s["hypo"][:, 1], stats.norm(syn.y_mu[0], cfg.prior_width_xy).cdf, thin)
    assert ks_ok(s["hypo"][:, 2] - cfg.prior_z, stats.rayleigh(scale=cfg.prior_width_z).cdf, thin)
    assert ks_ok(s["vs"], stats.norm(cfg.prior_vs, cfg.prior_width_vs).cdf, 400)
    assert ks_ok(s["qs"], stats.norm(cfg.prior_qs, cfg.prior_width_qs).cdf, 400)
    assert ks_ok(s["t_corr"][:, 2], stats.norm(cfg.prior_t_corr, cfg.prior_width_t_corr).cdf, 1500)
    assert ks_ok(s["a_corr"][:, 4], stats.norm(cfg.prior_a_corr, cfg.prior_width_a_corr).cdf, 1500)


def test_prior_only_factorised_schedule():
    syn = H.Synthetic(3, 6, 22)
    cfg = H.default_config(n_sta=6, n_events=3, n_procs=2, n_chains=4, n_cool=1, n_iter=60000, n_burn=500,
                           n_interval=5, mode=H.MODE_FACTORISED, precision=64, use_time=0, use_amp=0,
                           step_size_xy=45.0, step_size_z=12.0, **NOSOLVE)
    o = Oracle(cfg, syn)
    o.init_chains()
    o.run(1, cfg.n_iter, trace=False)
    s = cold_samples(o, 2)
    for e in range(3):
        assert ks_ok(s["hypo"][:, 3 * e], stats.norm(syn.x_mu[e], cfg.prior_width_xy).cdf, 10)
        assert ks_ok(s["hypo"][:, 3 * e + 1], stats.norm(syn.y_mu[e], cfg.prior_width_xy).cdf, 10)
        assert ks_ok(s["hypo"][:, 3 * e + 2] - cfg.prior_z, stats.rayleigh(scale=cfg.prior_width_z).cdf, 10)


def test_factorised_and_reference_schedules_share_the_posterior():
    # shared parameters fixed => the posterior factorises over events => mode B (per-event tempered
    # chains) and mode A (joint chain) must agree on every cold-chain marginal
    syn = H.Synthetic(2, 8, 23)
    base = dict(n_sta=8, n_events=2, n_procs=2, n_chains=4, n_cool=1, n_burn=5000, precision=64, **NOSOLVE)
    cfgA = H.default_config(mode=H.MODE_REPLAY, n_iter=600000, n_interval=11, **base)
    cfgB = H.default_config(mode=H.MODE_FACTORISED, n_iter=200000, n_interval=7, **base)
    a, b = Oracle(cfgA, syn), Oracle(cfgB, syn)
    a.init_chains()
    b.init_chains()
    a.run(1, cfgA.n_iter, trace=False)
    b.run(1, cfgB.n_iter, trace=False)
    sa, sb = cold_samples(a, 2)["hypo"], cold_samples(b, 2)["hypo"]
    assert len(sa) > 50000 and len(sb) > 50000
    for c in range(6):
        xa, xb = sa[::60, c], sb[::30, c]
        assert stats.ks_2samp(xa, xb).pvalue > 1e-3, "marginal %d differs" % c
        # medians agree within a few Monte-Carlo standard errors of the posterior spread
        spread = np.std(xa)
        assert abs(np.median(xa) - np.median(xb)) < 0.15 * spread
    # and both sit near the true hypocentres
    truth = np.stack([syn.true_x, syn.true_y, syn.true_z], axis=1).ravel()
    assert np.all(np.abs(np.median(sb, axis=0) - truth) < 4 * np.std(sb, axis=0) + 0.5)


def test_prior_only_blocked_gibbs_schedule():
    # mode C (every event proposes per iteration, then one shared parameter judged on the sum over events) with no
    # data: every cold marginal -- hypocentres, vs, qs, station terms -- is its prior
    syn = H.Synthetic(2, 6, 24)
    cfg = H.default_config(n_sta=6, n_events=2, n_procs=2, n_chains=3, n_cool=1, n_iter=150000, n_burn=1000,
                           n_interval=5, mode=H.MODE_BLOCKED_GIBBS, precision=64, use_time=0, use_amp=0,
                           step_size_xy=45.0, step_size_z=12.0, step_size_vs=1.5, step_size_qs=150.0,
                           step_size_t_corr=0.8, step_size_a_corr=0.03)
    o = Oracle(cfg, syn)
    o.init_chains()
    o.run(1, cfg.n_iter, trace=False)
    s = cold_samples(o, 2)
    assert len(s["vs"]) > 50000
    for e in range(2):
        assert ks_ok(s["hypo"][:, 3 * e], stats.norm(syn.x_mu[e], cfg.prior_width_xy).cdf, 20)
        assert ks_ok(s["hypo"][:, 3 * e + 1], stats.norm(syn.y_mu[e], cfg.prior_width_xy).cdf, 20)
        assert ks_ok(s["hypo"][:, 3 * e + 2] - cfg.prior_z, stats.rayleigh(scale=cfg.prior_width_z).cdf, 20)
    # one shared-parameter proposal per iteration, 14 targets (vs, qs, 6 + 6 station terms): slower mixing
    assert ks_ok(s["vs"], stats.norm(cfg.prior_vs, cfg.prior_width_vs).cdf, 60)
    assert ks_ok(s["qs"], stats.norm(cfg.prior_qs, cfg.prior_width_qs).cdf, 60)
    assert ks_ok(s["t_corr"][:, 2], stats.norm(cfg.prior_t_corr, cfg.prior_width_t_corr).cdf, 300)
    assert ks_ok(s["a_corr"][:, 4], stats.norm(cfg.prior_a_corr, cfg.prior_width_a_corr).cdf, 300)


def test_blocked_gibbs_and_reference_schedules_share_the_posterior():
    # shared parameters SOLVED: the joint posterior does not factorise any more; the blocked-Gibbs schedule (mode C)
    # must still have the reference schedule's (mode A) cold-chain marginals for hypocentres, vs and qs
    syn = H.Synthetic(2, 8, 25)
    base = dict(n_sta=8, n_events=2, n_procs=2, n_chains=4, n_cool=1, n_burn=20000, precision=64,
                solve_vs=1, solve_qs=1, solve_t_corr=0, solve_a_corr=0)
    cfgA = H.default_config(mode=H.MODE_REPLAY, n_iter=1200000, n_interval=11, **base)
    cfgC = H.default_config(mode=H.MODE_BLOCKED_GIBBS, n_iter=300000, n_interval=5, **base)
    a, c = Oracle(cfgA, syn), Oracle(cfgC, syn)
    a.init_chains()
    c.init_chains()
    a.run_threaded(1, cfgA.n_iter)
    c.run(1, cfgC.n_iter, trace=False)
    sa, sc = cold_samples(a, 2), cold_samples(c, 2)
    assert len(sa["vs"]) > 50000 and len(sc["vs"]) > 50000
    for col in range(6):
        xa, xc = sa["hypo"][::150, col], sc["hypo"][::80, col]
        assert stats.ks_2samp(xa, xc).pvalue > 1e-3, "hypocentre marginal %d differs" % col
        assert abs(np.median(xa) - np.median(xc)) < 0.2 * np.std(xa)
    for k in ("vs", "qs"):
        xa, xc = sa[k][::400], sc[k][::300]
        assert stats.ks_2samp(xa, xc).pvalue > 1e-3, k
        assert abs(np.median(xa) - np.median(xc)) < 0.25 * np.std(xa), k
