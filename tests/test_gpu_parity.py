"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs ("Tolerances"), plus size-independent properties at
BASELINE.json's full sizes.

Tolerances (stated here because the reference states none):
  loglik float64   |L_gpu - L_oracle| <= 1e-12 * max(1, |L|) per event
  loglik float32   <= 2e-3 absolute per event at S = 50 (measured ~3e-4)
  replay (mode A)  proposal type / index / prior_ok / accept flags and swap records IDENTICAL at
                   every step; |L_gpu - L_oracle| <= 1e-9 * max(1, |L|)
  factorised f64   same, against the oracle's statement of the factorised schedule
  factorised f32   statistical: two-sample KS against float64 oracle samples (SURVEY tolerances at BASELINE sizes:
                   tests/test_gpu_posterior.py)
  blocked Gibbs    f64: step-exact like the others.  f32 (moment form, htm_gibbs_f32.cu): first iteration follows the
                   f64 kernel (per-event |dL| <= 2e-3 + 3e-5 |L|, > 99.8 % equal flags); the Metropolis difference of a
                   shared parameter at 20 000 events within 1e-2 log-likelihood units of a float64 recomputation;
                   carried sums within 2e-5 relative; posterior: tests/test_gpu_posterior.py
"""
import numpy as np
import pytest
from scipy import stats

import hypotremormcmc_b200 as H
from oracle.pyoracle import Oracle

pytestmark = pytest.mark.gpu

NOSOLVE = dict(solve_vs=0, solve_t_corr=0, solve_qs=0, solve_a_corr=0)
FLAGS = ("proposal_type", "index", "prior_ok", "accepted")


def rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))


def models(syn, rng, M):
    E, S = syn.n_events, syn.n_sta
    hypo = np.stack([np.stack([syn.true_x + rng.normal(0, 3, E), syn.true_y + rng.normal(0, 3, E),
                               syn.true_z + rng.normal(0, 1, E)], axis=1).ravel() for _ in range(M)])
    return (hypo, rng.normal(0, 0.2, (M, S)), rng.normal(0, 0.02, (M, S)), rng.uniform(2.6, 3.4, M),
            rng.uniform(180, 320, M))


# ---- cls_forward ---------------------------------------------------------------------------------
@pytest.mark.parametrize("E,S", [(1, 10), (37, 20), (100, 50), (5, 33), (3, 1), (2, 97)])
def test_loglik_float64(E, S):
    syn = H.Synthetic(E, S, 100 + E)
    args = models(syn, np.random.default_rng(E), 5)
    cfg = H.default_config(n_sta=S, n_events=E, mode=H.MODE_REPLAY, precision=64)
    Lo, po_ = Oracle(cfg, syn).loglik(*args, per_event=True)
    with H.HypoTremorB200(cfg) as g:
        g.load(syn)
        Lg, pg = g.loglik(*args, per_event=True)
    assert rel(pg, po_) <= 1e-12
    assert rel(Lg, Lo) <= 1e-12


def test_loglik_float32_tolerance():
    syn = H.Synthetic(200, 50, 7)
    args = models(syn, np.random.default_rng(3), 4)
    cfg = H.default_config(n_sta=50, n_events=200, mode=H.MODE_FACTORISED, precision=32, **NOSOLVE)
    Lo, po_ = Oracle(cfg, syn).loglik(*args, per_event=True)
    with H.HypoTremorB200(cfg) as g:
        g.load(syn)
        Lg, pg = g.loglik(*args, per_event=True)
    assert np.max(np.abs(pg - po_)) <= 2e-3
    assert rel(Lg, Lo) <= 1e-5


@pytest.mark.parametrize("precision", [64, 32])
@pytest.mark.parametrize("E,S,M", [(1, 10, 1), (77, 13, 2), (200, 50, 3), (95, 20, 5), (64, 7, 8), (131, 33, 9), (40, 1, 4)])
def test_loglik_tile_and_warp_kernels_agree_with_oracle(monkeypatch, precision, E, S, M):
    """htm_loglik has two kernels (32-event tiles with TMA-staged rows; one warp per (model, event)): both against
    the oracle, for every models-per-CTA / station-slice split of the tile kernel and ragged tiles."""
    syn = H.Synthetic(E, S, 11)
    args = models(syn, np.random.default_rng(5), M)
    cfg = H.default_config(n_sta=S, n_events=E, mode=H.MODE_FACTORISED, precision=precision, **NOSOLVE)
    Lo, po_ = Oracle(cfg, syn).loglik(*args, per_event=True)
    out = {}
    for kern in ("tile", "warp"):
        monkeypatch.setenv("HTM_LOGLIK_KERNEL", kern)
        with H.HypoTremorB200(cfg) as g:
            g.load(syn)
            out[kern] = g.loglik(*args, per_event=True)
    for kern, (Lg, pg) in out.items():
        if precision == 64:
            assert np.max(np.abs(pg - po_) / np.maximum(1.0, np.abs(po_))) <= 1e-12, kern
            assert rel(Lg, Lo) <= 1e-12, kern
        else:
            assert np.max(np.abs(pg - po_)) <= 2e-3 and rel(Lg, Lo) <= 1e-5, kern


def test_loglik_golden_and_degenerate_sigma():
    import os
    import types
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "forward_golden.npz"))
    s = types.SimpleNamespace()
    s.sta_x, s.sta_y, s.sta_z = G["sta"]
    s.t_obs, s.t_stdv, s.a_obs, s.a_stdv = G["t_obs"], G["t_stdv"], G["a_obs"], G["a_stdv"]
    s.n_events, s.n_sta = s.t_obs.shape
    s.x_mu, s.y_mu = np.zeros(s.n_events), np.zeros(s.n_events)
    for tag, ut, ua in (("both", 1, 1), ("time", 1, 0), ("amp", 0, 1)):
        cfg = H.default_config(n_sta=s.n_sta, n_events=s.n_events, mode=H.MODE_REPLAY, precision=64, use_time=ut,
                               use_amp=ua)
        with H.HypoTremorB200(cfg) as g:
            g.load(s)
            L, pe = g.loglik(G["hypo"], G["t_corr"], G["a_corr"], G["vs"], G["qs"], per_event=True)
        assert np.allclose(L, G["L_" + tag], rtol=1e-12)
        assert np.allclose(pe, G["per_event_" + tag], rtol=1e-12, atol=1e-12)


def test_loglik_linearity_property_full_size():
    # size-independent property at C3 size: shifting every observation of an event by a constant
    # leaves L unchanged (weighted demean, src/cls_forward.f90:166-173)
    syn = H.Synthetic(10000, 50, 8)
    args = models(syn, np.random.default_rng(4), 1)
    cfg = H.default_config(n_sta=50, n_events=10000, mode=H.MODE_REPLAY, precision=64)
    with H.HypoTremorB200(cfg) as g:
        g.load(syn)
        L0, p0 = g.loglik(*args, per_event=True)
        rng = np.random.default_rng(5)
        g.set_observations(syn.t_obs + rng.normal(0, 5, (10000, 1)), syn.t_stdv,
                           syn.a_obs + rng.normal(0, 2, (10000, 1)), syn.a_stdv)
        L1, p1 = g.loglik(*args, per_event=True)
    assert rel(p1, p0) <= 1e-10
    assert abs(L0[0] - p0.sum()) <= 1e-9 * abs(L0[0])


# ---- mode A replay (BASELINE config 5) --------------------------------------------------------------
def replay_case(E, S, R, K, n_it, seed, **kw):
    syn = H.Synthetic(E, S, seed)
    cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=n_it, n_burn=n_it // 4,
                           n_interval=10, mode=H.MODE_REPLAY, precision=64, **kw)
    o = Oracle(cfg, syn)
    o.init_chains()
    st = [[o.get_chain_state(r, j) for j in range(K)] for r in range(R)]
    o.record_draws(True)
    tr_o, sw_o = o.run(1, n_it)
    draws = [o.draws(r) for r in range(R)]
    with H.HypoTremorB200(cfg) as g:
        g.load(syn)
        for r in range(R):
            for j in range(K):
                s = st[r][j]
                g.set_chain_state(r, j, s["hypo"], s["t_corr"], s["a_corr"], s["vs"], s["qs"], s["temp"],
                                  s["log_likelihood"])
        tr_g, sw_g, used = g.replay(1, n_it, draws)
        cg = g.get_counts()
        fin = [[g.get_chain_state(r, j) for j in range(K)] for r in range(R)]
    for f in FLAGS:
        assert np.array_equal(tr_o[f], tr_g[f]), f
    assert np.array_equal(sw_o, sw_g)
    assert rel(tr_g["log_likelihood"], tr_o["log_likelihood"]) <= 1e-9
    assert np.array_equal(used, [len(d) for d in draws])
    co = o.get_counts()
    assert np.array_equal(co[0], cg[0]) and np.array_equal(co[1], cg[1])
    for r in range(R):
        for j in range(K):
            so = o.get_chain_state(r, j)
            assert np.allclose(fin[r][j]["hypo"], so["hypo"], rtol=1e-11, atol=1e-11)
            assert fin[r][j]["temp"] == so["temp"]
            assert abs(fin[r][j]["vs"] - so["vs"]) < 1e-12 and np.allclose(fin[r][j]["t_corr"], so["t_corr"], atol=1e-12)


def test_replay_config5_100_events_50_stations():
    replay_case(100, 50, 4, 5, 3000, 20231006)


def test_replay_config1_single_event():
    replay_case(1, 10, 4, 5, 4000, 20231002)


def test_replay_fixed_globals_and_many_ranks():
    replay_case(6, 7, 11, 2, 600, 31, **NOSOLVE)


def test_replay_reports_exhausted_draws():
    syn = H.Synthetic(3, 6, 1)
    cfg = H.default_config(n_sta=6, n_events=3, n_procs=2, n_chains=2, n_iter=50, n_burn=0, n_interval=10,
                           mode=H.MODE_REPLAY, precision=64)
    o = Oracle(cfg, syn)
    o.init_chains()
    o.record_draws(True)
    o.run(1, 50)
    with H.HypoTremorB200(cfg) as g:
        g.load(syn)
        for r in range(2):
            for j in range(2):
                g.set_chain_state(r, j, np.ones(9) * 5, np.zeros(6), np.zeros(6), 3.0, 250.0, 1.0, -9e300)
        with pytest.raises(H.HtmError) as ei:
            g.replay(1, 50, [o.draws(0)[:40], o.draws(1)[:40]])
        assert ei.value.code == H.config.HTM_ERR_DRAWS


# ---- mode B factorised: step-exact in float64 against the oracle's statement of the schedule --------
@pytest.mark.parametrize("kernel,slots", [(2, 1), (2, 2), (2, 4), (1, 0)])
@pytest.mark.parametrize("E,S,R,K", [(6, 10, 4, 5), (5, 50, 2, 16), (3, 20, 3, 1), (4, 33, 5, 7), (2, 12, 1, 32),
                                     (1, 1, 1, 2), (3, 130, 2, 3), (7, 5, 9, 4)])
def test_factorised_float64_step_exact(kernel, slots, E, S, R, K):
    if kernel == 1 and (K > 16 or S > 128):
        pytest.skip("warp-per-chain kernel: n_chains <= 16, n_sta <= 128")
    syn = H.Synthetic(E, S, 13 + E)
    cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=80, n_burn=20, n_interval=8,
                           mode=H.MODE_FACTORISED, precision=64, kernel=kernel, lane_slots=slots, max_samples=16,
                           hist_bins=16, **NOSOLVE)
    o = Oracle(cfg, syn)
    o.init_chains()
    tr_o, sw_o = o.run(1, 80)
    with H.HypoTremorB200(cfg) as g:
        g.load(syn)
        g.init_chains()
        tr_g, sw_g = g.run_traced(1, 80)
        cg = g.get_counts()
        smp = [g.fetch_samples(r) for r in range(R)]
        lik = [g.fetch_likelihood(r) for r in range(R)]
        hist = g.get_histograms()
    for f in FLAGS:
        assert np.array_equal(tr_o[f], tr_g[f]), f
    if K >= 2:
        assert np.array_equal(sw_o, sw_g)
    assert rel(tr_g["log_likelihood"], tr_o["log_likelihood"]) <= 1e-9
    co = o.get_counts()
    assert np.array_equal(co[0], cg[0]) and np.array_equal(co[1], cg[1])
    for r in range(R):
        so = o.fetch_samples(r)
        assert np.array_equal(so["iter"], smp[r]["iter"])
        assert np.allclose(so["hypo"], smp[r]["hypo"], rtol=1e-10, atol=1e-10)
        lo = o.fetch_likelihood(r)
        assert np.array_equal(lo[0], lik[r][0]) and np.allclose(lo[1], lik[r][1], rtol=1e-10)
    # every post-burn-in cold sample landed in exactly one bin per coordinate
    n_rec_post = sum(len(s["iter"]) for s in smp)
    assert hist.sum() == 3 * n_rec_post * E


@pytest.mark.parametrize("E,S,R,K,n_cool", [(3, 9, 2, 33, 1), (2, 20, 1, 70, 3), (4, 7, 3, 128, 2)])
def test_factorised_wide_groups_float64_step_exact(E, S, R, K, n_cool):
    """Tempering groups of more than 32 chains (the reference has no limit on n_chains): one CTA per group, swap and
    record numbering through shared memory (fact_wide_kernel) -- step-exact against the oracle like the others."""
    syn = H.Synthetic(E, S, 90 + K)
    cfg = fact_cfg(E, S, R, K, precision=64, n_iter=60, n_interval=6, n_burn=12, n_cool=n_cool, max_samples=16, hist_bins=8)
    o = Oracle(cfg, syn)
    o.init_chains()
    tr_o, sw_o = o.run(1, 60)
    with H.HypoTremorB200(cfg) as g:
        g.load(syn)
        g.init_chains()
        tr_g, sw_g = g.run_traced(1, 35)
        tr_g2, sw_g2 = g.run_traced(36, 60)
        cg = g.get_counts()
        smp = [g.fetch_samples(r) for r in range(R)]
        lik = [g.fetch_likelihood(r) for r in range(R)]
    tr_g, sw_g = np.concatenate([tr_g, tr_g2]), np.concatenate([sw_g, sw_g2])
    for f in FLAGS:
        assert np.array_equal(tr_o[f], tr_g[f]), f
    assert np.array_equal(sw_o, sw_g) and rel(tr_g["log_likelihood"], tr_o["log_likelihood"]) <= 1e-9
    co = o.get_counts()
    assert np.array_equal(co[0], cg[0]) and np.array_equal(co[1], cg[1])
    for r in range(R):
        so = o.fetch_samples(r)
        assert np.array_equal(so["iter"], smp[r]["iter"]) and np.allclose(so["hypo"], smp[r]["hypo"], rtol=1e-10, atol=1e-10)
        lo = o.fetch_likelihood(r)
        assert np.array_equal(lo[0], lik[r][0]) and np.allclose(lo[1], lik[r][1], rtol=1e-10)
    # float32: runs, counts every cold proposal, keeps the temperature multiset
    c32 = H.copy_config(cfg, precision=32, max_samples=64)
    with H.HypoTremorB200(c32) as g:
        g.load(syn)
        g.init_chains()
        t0 = sorted(g.get_chain_state(0, k)["temp"] for k in range(K))
        g.run(1, 300)
        p, a = g.get_counts()
        t1 = sorted(g.get_chain_state(0, k)["temp"] for k in range(K))
    assert p[4:].sum() == 300 * E * R * n_cool and 0 < a.sum() < p.sum() and t0 == t1


def test_factorised_two_cold_chains_and_geometric_ladder():
    syn = H.Synthetic(4, 15, 3)
    cfg = H.default_config(n_sta=15, n_events=4, n_procs=2, n_chains=6, n_cool=2, n_iter=60, n_burn=0, n_interval=6,
                           mode=H.MODE_FACTORISED, precision=64, ladder=H.LADDER_GEOMETRIC, max_samples=16, **NOSOLVE)
    o = Oracle(cfg, syn)
    o.init_chains()
    tr_o, sw_o = o.run(1, 60)
    for kernel in (1, 2):
        c = H.copy_config(cfg, kernel=kernel)
        with H.HypoTremorB200(c) as g:
            g.load(syn)
            g.init_chains()
            tr_g, sw_g = g.run_traced(1, 60)
            s0 = g.fetch_samples(0)
        for f in FLAGS:
            assert np.array_equal(tr_o[f], tr_g[f])
        assert np.array_equal(sw_o, sw_g)
    so = o.fetch_samples(0)
    assert len(so["iter"]) == 20 and np.allclose(so["hypo"], s0["hypo"], rtol=1e-10, atol=1e-10)


def test_factorised_fixed_globals_are_folded_exactly():
    syn = H.Synthetic(3, 9, 5)
    rng = np.random.default_rng(6)
    tc, ac = rng.normal(0, 0.2, 9), rng.normal(0, 0.02, 9)
    cfg = H.default_config(n_sta=9, n_events=3, n_procs=1, n_chains=4, n_iter=50, n_burn=0, n_interval=5,
                           mode=H.MODE_FACTORISED, precision=64, **NOSOLVE)
    o = Oracle(cfg, syn)
    o.set_globals(2.7, 310.0, tc, ac)
    o.init_chains()
    tr_o, _ = o.run(1, 50)
    with H.HypoTremorB200(cfg) as g:
        g.load(syn)
        g.set_globals(2.7, 310.0, tc, ac)
        g.init_chains()
        tr_g, _ = g.run_traced(1, 50)
    for f in FLAGS:
        assert np.array_equal(tr_o[f], tr_g[f])
    assert rel(tr_g["log_likelihood"], tr_o["log_likelihood"]) <= 1e-9


@pytest.mark.parametrize("kernel,slots", [(2, 1), (2, 2), (2, 4), (1, 0)])
@pytest.mark.parametrize("S", [20, 50, 33, 7, 1])
def test_factorised_float32_kernel_likelihood_accuracy(kernel, slots, S):
    # the float32 throughput kernels' own log-likelihood (carried in registers across steps) against a
    # float64 evaluation of the SAME float32 states by the oracle: isolates the kernels' arithmetic
    # (distance by expansion, sqrt-weight folding, MUFU approximations).  Bound per event:
    # 2e-3 + 5e-5 |L|  (measured: <= 1.5e-3 for |L| < 500, <= 2.5e-5 |L| for far-from-converged hot chains).
    E, R, K = 24, 4, 8
    syn = H.Synthetic(E, S, 17)
    cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=40, n_burn=0, n_interval=10,
                           mode=H.MODE_FACTORISED, precision=32, kernel=kernel, lane_slots=slots, **NOSOLVE)
    o = Oracle(cfg, syn)
    with H.HypoTremorB200(cfg) as g:
        g.load(syn)
        g.init_chains()
        tr, _ = g.run_traced(1, 40)
        worst = 0.0
        for r in range(R):
            for k in range(K):
                st = g.get_chain_state(r, k)
                _, pe = o.loglik(st["hypo"][None, :], np.zeros((1, S)), np.zeros((1, S)), [cfg.prior_vs],
                                 [cfg.prior_qs], per_event=True)
                err = np.abs(pe[0] - tr["log_likelihood"][-1, :, r, k])
                worst = max(worst, np.max(err / (2e-3 + 5e-5 * np.abs(pe[0]))))
    assert worst <= 1.0, worst


# ---- size-independent properties at BASELINE sizes ------------------------------------------------------
def fact_cfg(E, S, R, K, **kw):
    base = dict(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=400, n_burn=0, n_interval=50,
                mode=H.MODE_FACTORISED, precision=32, **NOSOLVE)
    base.update(kw)
    return H.default_config(**base)


def final_state(g, R, K):
    return [g.get_chain_state(r, k) for r in range(R) for k in (0, K - 1)]


def test_chunked_runs_equal_one_run_config2():
    # idempotence of the launch boundary: run(1,400) == run(1,150) + run(151,400), bit for bit
    syn = H.Synthetic(1000, 20, 20231002)
    cfg = fact_cfg(1000, 20, 4, 16, hist_bins=32)
    with H.HypoTremorB200(cfg) as a, H.HypoTremorB200(cfg) as b:
        for g in (a, b):
            g.load(syn)
            g.init_chains()
        a.run(1, 400)
        b.run(1, 150)
        b.run(151, 400)
        sa, sb = final_state(a, 4, 16), final_state(b, 4, 16)
        for x, y in zip(sa, sb):
            assert np.array_equal(x["hypo"], y["hypo"]) and x["log_likelihood"] == y["log_likelihood"]
        assert np.array_equal(a.get_histograms(), b.get_histograms())
        assert np.array_equal(a.get_counts()[0], b.get_counts()[0])


def test_kernel_layouts_agree_config2():
    # lane-per-chain (1, 2, 4 slots) and warp-per-chain run the same chains: identical accept counts in
    # float64; in float32 the summation order differs, so compare statistically
    syn = H.Synthetic(500, 20, 77)
    counts = []
    for kernel, slots in ((2, 1), (2, 2), (2, 4), (1, 0)):
        cfg = fact_cfg(500, 20, 4, 16, precision=64, kernel=kernel, lane_slots=slots, n_iter=100)
        with H.HypoTremorB200(cfg) as g:
            g.load(syn)
            g.init_chains()
            g.run(1, 100)
            counts.append(g.get_counts())
    for c in counts[1:]:
        assert np.array_equal(c[0], counts[0][0])
        assert np.array_equal(c[1], counts[0][1])


def test_event_sharding_is_invariant_config2():
    # events shard with no exchange step: a shard reproduces its slice of the unsharded run exactly
    syn = H.Synthetic(1000, 20, 20231002)
    full_cfg = fact_cfg(1000, 20, 4, 16, n_iter=200)
    with H.HypoTremorB200(full_cfg) as g:
        g.load(syn)
        g.init_chains()
        g.run(1, 200)
        full = g.get_chain_state(2, 0)
        full_counts = g.get_counts()
    tot = [np.zeros(7, dtype=np.int64), np.zeros(7, dtype=np.int64)]
    for rank in range(3):
        sh = syn.shard(rank, 3)
        cfg = H.copy_config(full_cfg, shard_rank=rank, shard_count=3)
        with H.HypoTremorB200(cfg) as g:
            g.load(sh)
            g.init_chains()
            g.run(1, 200)
            part = g.get_chain_state(2, 0)
            c = g.get_counts()
        lo = sh.event_offset
        assert np.array_equal(part["hypo"], full["hypo"][3 * lo:3 * (lo + sh.n_events)])
        tot[0] += c[0]
        tot[1] += c[1]
    assert np.array_equal(tot[0], full_counts[0]) and np.array_equal(tot[1], full_counts[1])


def test_counts_and_temperatures_conserved_100k_chains():
    # >= 100k tempered chains of a 50-station network in lockstep (north_star target size)
    E, S, R, K = 2000, 50, 4, 16      # 128,000 chains
    syn = H.Synthetic(E, S, 9)
    cfg = fact_cfg(E, S, R, K, n_iter=300, n_interval=10, hist_bins=16)
    with H.HypoTremorB200(cfg) as g:
        g.load(syn)
        g.init_chains()
        g.run(1, 300)
        p, a = g.get_counts()
        hist = g.get_histograms()
    assert p[:4].sum() == 0 and p[4:].sum() == 300 * E * R      # exactly one cold chain per group, every step
    assert np.all(a <= p) and 0.05 < a.sum() / p.sum() < 0.95
    assert hist.sum() == 3 * 30 * E * R


# ---- float32 throughput path: posterior marginals against the float64 reference-schedule oracle -----------
def test_float32_posterior_matches_reference_schedule_oracle():
    # correctness part (2) of the north star: posterior marginals of x, y, depth from the B200 float32
    # factorised kernel vs the oracle's reference schedule (mode A, joint chain, mod_random).
    # Pass: two-sample KS p > 1e-3 on thinned samples for every marginal, medians within 0.15 sigma.
    syn = H.Synthetic(2, 8, 23)
    base = dict(n_sta=8, n_events=2, n_procs=2, n_chains=4, n_cool=1, n_burn=5000, **NOSOLVE)
    cfgA = H.default_config(mode=H.MODE_REPLAY, precision=64, n_iter=600000, n_interval=11, **base)
    o = Oracle(cfgA, syn)
    o.init_chains()
    o.run(1, cfgA.n_iter, trace=False)
    sa = np.concatenate([o.fetch_samples(r)["hypo"] for r in range(2)])
    n_it, interval = 200000, 7
    for kernel in (1, 2):
        cfgB = H.default_config(mode=H.MODE_FACTORISED, precision=32, n_iter=n_it, n_interval=interval, kernel=kernel,
                                max_samples=n_it // interval + 2, **base)
        with H.HypoTremorB200(cfgB) as g:
            g.load(syn)
            g.init_chains()
            g.run(1, n_it)
            sb = np.concatenate([g.fetch_samples(r)["hypo"] for r in range(2)])
        assert len(sb) > 50000
        for c in range(6):
            xa, xb = sa[::60, c], sb[::30, c]
            assert stats.ks_2samp(xa, xb).pvalue > 1e-3, "kernel %d marginal %d" % (kernel, c)
            assert abs(np.median(xa) - np.median(xb)) < 0.15 * np.std(xa)
            qa, qb = np.quantile(xa, [0.025, 0.975]), np.quantile(xb, [0.025, 0.975])
            assert np.all(np.abs(qa - qb) < 0.35 * np.std(xa))


def test_float32_posterior_at_config2_size_matches_float64_oracle_on_a_subset():
    # BASELINE configs[1] at full size on the GPU (1000 events x 20 stations x 16 temperatures x 4 chains, float32)
    # against the float64 oracle run on a SUBSET of the same events (events are independent in this mode, and
    # Philox ids are global, so the oracle samples exactly those events' chains).  KS p > 1e-3 per marginal,
    # medians within 0.15 sigma.
    E, S, R, K = 1000, 20, 4, 16
    n_it, burn, interval = 12000, 2000, 5
    syn = H.Synthetic(E, S, 20231002)
    cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=n_it, n_burn=burn,
                           n_interval=interval, mode=H.MODE_FACTORISED, precision=32, seed=77,
                           max_samples=n_it // interval + 2, **NOSOLVE)
    with H.HypoTremorB200(cfg) as g:
        g.load(syn)
        g.init_chains()
        g.run(1, n_it)
        sg = np.concatenate([g.fetch_samples(r)["hypo"] for r in range(R)])
    assert sg.shape == (R * (n_it - burn) // interval, 3 * E)
    lo, n_sub = 417, 6                                    # events 417..422
    sub = syn.shard(0, 1)
    for name in ("true_x", "true_y", "true_z", "x_mu", "y_mu", "t_obs", "t_stdv", "a_obs", "a_stdv"):
        setattr(sub, name, np.ascontiguousarray(getattr(syn, name)[lo:lo + n_sub]))
    sub.n_events = n_sub
    ocfg = H.copy_config(cfg, precision=64, n_events=n_sub)
    o = Oracle(ocfg, sub, event_offset=lo)
    o.init_chains()
    o.run(1, n_it, trace=False)
    so = np.concatenate([o.fetch_samples(r)["hypo"] for r in range(R)])
    for e in range(n_sub):
        for c in range(3):
            a, b = so[::8, 3 * e + c], sg[::8, 3 * (lo + e) + c]
            assert stats.ks_2samp(a, b).pvalue > 1e-3, (e, c)
            assert abs(np.median(a) - np.median(b)) < 0.15 * np.std(a), (e, c)


def test_float32_prior_only_samples_the_priors():
    syn = H.Synthetic(64, 6, 22)
    cfg = H.default_config(n_sta=6, n_events=64, n_procs=2, n_chains=4, n_cool=1, n_iter=20000, n_burn=500,
                           n_interval=5, mode=H.MODE_FACTORISED, precision=32, use_time=0, use_amp=0,
                           step_size_xy=45.0, step_size_z=12.0, max_samples=4002, **NOSOLVE)
    with H.HypoTremorB200(cfg) as g:
        g.load(syn)
        g.init_chains()
        g.run(1, 20000)
        s = np.concatenate([g.fetch_samples(r)["hypo"] for r in range(2)])
    x = (s[::10, 0::3] - syn.x_mu[None, :]).ravel()
    z = (s[::10, 2::3] - cfg.prior_z).ravel()
    assert stats.kstest(x, stats.norm(0, cfg.prior_width_xy).cdf).pvalue > 1e-3
    assert stats.kstest(z, stats.rayleigh(scale=cfg.prior_width_z).cdf).pvalue > 1e-3


def test_fp32_peak_microbenchmark_is_sane():
    tf, mufu = H.api.measure_fp32_peak(0)
    assert 30.0 < tf < 90.0          # nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4 TFLOP/s
    assert 2000.0 < mufu < 6000.0    # nominal 148 SM x 16 /clk x 1.965 GHz = 4653 Gop/s


def test_nccl_gather_of_histograms_and_counts():
    # the path's only collective (posterior-histogram gather), here over a 1-rank NCCL group
    import torch
    import torch.distributed as dist
    from hypotremormcmc_b200.gather import gather_run
    import os, socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        syn = H.Synthetic(300, 20, 5)
        cfg = fact_cfg(300, 20, 4, 16, n_iter=200, n_interval=10, hist_bins=32)
        with H.HypoTremorB200(cfg) as g:
            g.load(syn)
            g.init_chains()
            g.run(1, 200)
            hist, counts = gather_run(g)
            h_ref = g.get_histograms()
            p, a = g.get_counts()
        assert np.array_equal(hist.cpu().numpy().astype(np.uint32), h_ref)
        assert np.array_equal(counts.cpu().numpy(), np.concatenate([p, a]))
    finally:
        dist.destroy_process_group()


# ---- mode C blocked Gibbs (shared parameters solved) ---------------------------------------------------------
@pytest.fixture(params=[0, 1], ids=["launch-per-iteration", "persistent"])
def gibbs_path(request, monkeypatch):
    # HTM_GIBBS_PERSIST: 0 = one launch per iteration, 1 = the persistent cooperative kernel
    monkeypatch.setenv("HTM_GIBBS_PERSIST", str(request.param))
    return request.param


@pytest.mark.parametrize("E,S,R,K,solve", [(5, 9, 2, 3, (1, 1, 1, 1)), (40, 20, 3, 4, (1, 0, 1, 0)), (3, 33, 1, 2, (0, 1, 0, 1)),
                                            (70, 12, 5, 2, (1, 1, 1, 1)), (4, 8, 2, 3, (0, 0, 0, 0)),
                                            (1, 10, 4, 5, (1, 1, 1, 1))])  # the last: the shape of BASELINE configs[0]
def test_blocked_gibbs_float64_step_exact(gibbs_path, E, S, R, K, solve):
    syn = H.Synthetic(E, S, 40 + E)
    cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=60, n_burn=12, n_interval=6,
                           mode=H.MODE_BLOCKED_GIBBS, precision=64, max_samples=16, solve_vs=solve[0],
                           solve_t_corr=solve[1], solve_qs=solve[2], solve_a_corr=solve[3])
    o = Oracle(cfg, syn)
    o.init_chains()
    st_o = o.get_chain_state(R - 1, K - 1)
    tr_o, sw_o = o.run(1, 60)
    with H.HypoTremorB200(cfg) as g:
        g.load(syn)
        g.init_chains()
        st_g = g.get_chain_state(R - 1, K - 1)
        tr_g, sw_g = g.run_traced(1, 60)
        cg = g.get_counts()
        smp = [g.fetch_samples(r) for r in range(R)]
        lik = [g.fetch_likelihood(r) for r in range(R)]
        fin = g.get_chain_state(0, 0)
    assert np.allclose(st_g["hypo"], st_o["hypo"], rtol=1e-12, atol=1e-12) and st_g["temp"] == pytest.approx(st_o["temp"], rel=1e-13)
    assert np.allclose(st_g["t_corr"], st_o["t_corr"], atol=1e-14) and rel(st_g["log_likelihood"], st_o["log_likelihood"]) < 1e-12
    for f in FLAGS:
        assert np.array_equal(tr_o[f], tr_g[f]), f
    assert np.array_equal(sw_o, sw_g)
    assert rel(tr_g["log_likelihood"], tr_o["log_likelihood"]) <= 1e-9
    co = o.get_counts()
    assert np.array_equal(co[0], cg[0]) and np.array_equal(co[1], cg[1])
    for r in range(R):
        so = o.fetch_samples(r)
        assert np.array_equal(so["iter"], smp[r]["iter"])
        for k in ("hypo", "vs", "qs", "t_corr", "a_corr"):
            assert np.allclose(so[k], smp[r][k], rtol=1e-10, atol=1e-10), k
        lo = o.fetch_likelihood(r)
        assert np.array_equal(lo[0], lik[r][0]) and np.allclose(lo[1], lik[r][1], rtol=1e-10)
    fo = o.get_chain_state(0, 0)
    assert np.allclose(fin["hypo"], fo["hypo"], rtol=1e-10, atol=1e-10) and abs(fin["vs"] - fo["vs"]) < 1e-12
    assert rel(fin["log_likelihood"], fo["log_likelihood"]) < 1e-9


def test_blocked_gibbs_float32_first_iteration_follows_the_float64_kernel():
    """The float32 kernel (htm_gibbs_f32.cu: one pass per event + moment-based shared-parameter deltas) draws the same
    Philox words as the float64 kernels, so until rounding flips a decision the two runs coincide: at iteration 1
    every per-event log-likelihood and every chain's summed log-likelihood after the shared-parameter step must
    agree to float32 accuracy, and (almost) every accept flag must be the same."""
    E, S, R, K = 400, 20, 3, 4
    syn = H.Synthetic(E, S, 8)
    base = dict(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_iter=10, n_burn=0, n_interval=5,
                mode=H.MODE_BLOCKED_GIBBS, max_samples=4)
    tr = {}
    for prec in (64, 32):
        with H.HypoTremorB200(H.default_config(precision=prec, **base)) as g:
            g.load(syn)
            g.init_chains()
            tr[prec], _ = g.run_traced(1, 1)
    a, b = tr[64][0], tr[32][0]
    assert np.array_equal(a["proposal_type"], b["proposal_type"]) and np.array_equal(a["index"], b["index"])
    assert (a["accepted"] != b["accepted"]).mean() < 2e-3
    same = a["accepted"][:-1] == b["accepted"][:-1]
    dL = (np.abs(a["log_likelihood"][:-1] - b["log_likelihood"][:-1]) /
          (2e-3 + 3e-5 * np.abs(a["log_likelihood"][:-1])))[same]               # per event: abs + float32 relative
    assert dL.max() < 1.0, dL.max()
    tot = np.abs(a["log_likelihood"][-1] - b["log_likelihood"][-1])            # per chain, after the shared step
    agree = a["accepted"][-1] == b["accepted"][-1]
    assert agree.mean() > 0.8
    # (a flipped hypocentre decision somewhere moves the whole sum, so only a loose bound holds per chain)
    assert np.median(tot[agree] / np.abs(a["log_likelihood"][-1][agree])) < 1e-5


def test_blocked_gibbs_float32_shared_parameter_ratio_at_20000_events():
    """The Metropolis ratio of a shared parameter is a sum over ALL events.  The float32 kernel accumulates it in
    float64 from per-event DIFFERENCES (moment form, htm_gibbs_f32.cu); here the difference it judged is redone in
    float64 (htm_loglik on the state and on the state with the proposal applied) for real pending proposals of
    every kind at 20 000 events x 50 stations: |error| << 1 log-likelihood unit, however large the sums are."""
    E, S, R, K = 20000, 50, 2, 4
    J = R * K
    syn = H.Synthetic(E, S, 77)
    # small steps so that proposals of every kind stay in the range where they are sometimes accepted
    cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=100, n_burn=0, n_interval=50,
                           mode=H.MODE_BLOCKED_GIBBS, precision=32, step_size_vs=2e-4, step_size_qs=0.5,
                           step_size_t_corr=2e-3, step_size_a_corr=1e-3)
    worst, kinds = 0.0, set()
    # float64 evaluator of the same data (htm_loglik computes in the handle's precision)
    cfg64 = H.default_config(n_sta=S, n_events=E, mode=H.MODE_FACTORISED, precision=64, **NOSOLVE)
    with H.HypoTremorB200(cfg) as g, H.HypoTremorB200(cfg64) as g64:
        g.load(syn)
        g64.load(syn)
        g.init_chains()
        g.run(1, 30)
        it = 31
        for _ in range(6):
            before = [g.get_chain_state(c // K, c % K) for c in range(J)]
            which, idx, xn = g.gibbs_pending()
            g.run(it, it)
            cur32, prop32 = g.gibbs_last_sums()
            after = [g.get_chain_state(c // K, c % K) for c in range(J)]
            hypo = np.stack([s["hypo"] for s in after])           # hypocentres the shared step was judged on
            tc = np.stack([s["t_corr"] for s in before])
            ac = np.stack([s["a_corr"] for s in before])
            vs = np.array([s["vs"] for s in before])
            qs = np.array([s["qs"] for s in before])
            tcp, acp, vsp, qsp = tc.copy(), ac.copy(), vs.copy(), qs.copy()
            for c in range(J):
                kinds.add(int(which[c]))
                if which[c] == 1:
                    vsp[c] = xn[c]
                elif which[c] == 2:
                    tcp[c, idx[c]] = xn[c]
                elif which[c] == 3:
                    qsp[c] = xn[c]
                elif which[c] == 4:
                    acp[c, idx[c]] = xn[c]
            L_cur = g64.loglik(hypo, tc, ac, vs, qs)
            L_prop = g64.loglik(hypo, tcp, acp, vsp, qsp)
            err = np.abs((prop32 - cur32) - (L_prop - L_cur))
            worst = max(worst, float(err.max()))
            # the carried sums themselves are float32 accurate only: that is why differences are formed per event
            assert np.all(np.abs(cur32 - L_cur) <= 2e-5 * np.abs(L_cur))
            it += 1
    assert kinds >= {1, 2, 3, 4} or len(kinds) >= 3, kinds
    assert worst < 1e-2, worst


@pytest.mark.parametrize("scale", [1e-4, 1.0, 1e3])
def test_blocked_gibbs_float32_integer_sums_over_the_whole_range(scale):
    """The per-chain sums over events cross CTAs as two-limb fixed-point integers (htm_gibbs_f32.cu: publish_sum,
    2^-32 resolution).  With the observation errors scaled by 1e-4 a chain's sum is ~ -1e12 (1e22 units: far beyond one
    64-bit word, the high limb carries it), with 1e3 every event contributes ~ 1e-5: in both the judged sums must
    still equal a float64 evaluation of the same states to float32 accuracy."""
    E, S, R, K = 3000, 20, 2, 3
    J = R * K
    syn = H.Synthetic(E, S, 31)
    syn.t_stdv = syn.t_stdv * scale
    syn.a_stdv = syn.a_stdv * scale
    cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=20, n_burn=0, n_interval=10,
                           mode=H.MODE_BLOCKED_GIBBS, precision=32)
    cfg64 = H.default_config(n_sta=S, n_events=E, mode=H.MODE_FACTORISED, precision=64, **NOSOLVE)
    with H.HypoTremorB200(cfg) as g, H.HypoTremorB200(cfg64) as g64:
        g.load(syn)
        g64.load(syn)
        g.init_chains()
        g.run(1, 3)
        before = [g.get_chain_state(c // K, c % K) for c in range(J)]
        g.run(4, 4)
        cur32, prop32 = g.gibbs_last_sums()
        after = [g.get_chain_state(c // K, c % K) for c in range(J)]
        L_cur = g64.loglik(np.stack([s["hypo"] for s in after]), np.stack([s["t_corr"] for s in before]),
                           np.stack([s["a_corr"] for s in before]), [s["vs"] for s in before], [s["qs"] for s in before])
    assert np.all(np.isfinite(cur32)) and np.all(np.isfinite(prop32))
    if scale < 1.0:
        assert np.abs(L_cur).min() > 2.0 ** 31 * 100          # the sums do not fit one 64-bit word of 2^-32 units
    assert np.all(np.abs(cur32 - L_cur) <= 3e-5 * np.abs(L_cur)), (cur32, L_cur)


def test_blocked_gibbs_float32_many_joint_chains():
    """600 joint chains x 40 stations: the chain-level station terms (J x S x 2 doubles = 384 KB) no longer have to
    fit a CTA's shared memory -- the float32 kernel keeps them in global memory.  Size-independent invariants."""
    E, S, R, K, n_it = 500, 40, 100, 6, 40
    syn = H.Synthetic(E, S, 5)
    cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=n_it, n_burn=0, n_interval=10,
                           mode=H.MODE_BLOCKED_GIBBS, precision=32, max_samples=8)
    with H.HypoTremorB200(cfg) as g:
        g.load(syn)
        g.init_chains()
        temps0 = sorted(g.get_chain_state(r, k)["temp"] for r in range(0, R, 7) for k in range(K))
        g.run(1, n_it)
        p, a = g.get_counts()
        states = [g.get_chain_state(r, k) for r in range(0, R, 17) for k in range(K)]
        L64 = g.loglik(np.stack([s["hypo"] for s in states]), np.stack([s["t_corr"] for s in states]),
                       np.stack([s["a_corr"] for s in states]), [s["vs"] for s in states], [s["qs"] for s in states])
        smp = [g.fetch_samples(r) for r in range(R)]
    assert len(temps0) == len(range(0, R, 7)) * K
    assert p[4:7].sum() == n_it * E * R and p[:4].sum() == n_it * R and (a <= p).all()
    assert a[:4].sum() > 0 and a[4:7].sum() > 0
    for s, L in zip(states, L64):
        assert abs(s["log_likelihood"] - L) <= 2e-5 * abs(L), (s["log_likelihood"], L)
    assert sum(len(x["iter"]) for x in smp) == 4 * R      # iterations 1, 11, 21, 31: one cold chain per rank each
    with pytest.raises(H.HtmError) as ei:                  # the float64 kernels keep that state in shared memory
        with H.HypoTremorB200(H.copy_config(cfg, precision=64)) as g64:
            g64.load(syn)
            g64.init_chains()
            g64.run(1, 2)
    assert ei.value.code == H.config.HTM_ERR_UNSUPPORTED


def test_blocked_gibbs_full_size_properties():
    """20 000 events x 50 stations x 20 joint chains (float32 kernel, one cooperative launch): size-independent invariants."""
    E, S, R, K, n_it, n_int = 20000, 50, 4, 5, 120, 20
    syn = H.Synthetic(E, S, 77)
    cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=2, n_iter=n_it, n_burn=40, n_interval=n_int,
                           mode=H.MODE_BLOCKED_GIBBS, precision=32, max_samples=8)
    with H.HypoTremorB200(cfg) as g:
        g.load(syn)
        g.init_chains()
        temps0 = sorted(g.get_chain_state(r, k)["temp"] for r in range(R) for k in range(K))
        g.run(1, 70)
        g.run(71, n_it)
        _, nl, _ = g.last_run_stats()
        assert nl == 2  # prepare + one cooperative launch
        states = [g.get_chain_state(r, k) for r in range(R) for k in range(K)]
        p, a = g.get_counts()
        L64 = g.loglik(np.stack([s["hypo"] for s in states]), np.stack([s["t_corr"] for s in states]),
                       np.stack([s["a_corr"] for s in states]), [s["vs"] for s in states], [s["qs"] for s in states])
        smp = [g.fetch_samples(r) for r in range(R)]
        lik = [g.fetch_likelihood(r) for r in range(R)]
    # temperatures are only ever exchanged
    assert sorted(s["temp"] for s in states) == temps0
    # every cold chain proposes one hypocentre coordinate per event and one shared parameter per iteration
    n_cold = R * 2
    assert p[4:7].sum() == n_it * E * n_cold and p[:4].sum() == n_it * n_cold
    assert (a <= p).all() and 0.05 < a[4:7].sum() / p[4:7].sum() < 0.95
    # the carried log-likelihood of every chain is the likelihood of its state (float32 sums over 10^6 terms)
    for s, L in zip(states, L64):
        assert abs(s["log_likelihood"] - L) <= 2e-5 * abs(L), (s["log_likelihood"], L)
    # records: iterations 1, 21, ..., 101; hypocentres only after the burn-in (41, 61, 81, 101), n_cold per record
    n_rec = sum(len(x["iter"]) for x in smp)
    assert n_rec == 4 * n_cold and sum(len(l[0]) for l in lik) == 6 * n_cold
    for x in smp:
        assert set(np.unique(x["iter"])) <= {41, 61, 81, 101}
        assert np.isfinite(x["hypo"]).all() and (x["hypo"][:, 2::3] > cfg.prior_z).all()


def test_blocked_gibbs_float32_long_run_is_reproducible_and_consistent():
    """20 000 iterations of 100 joint chains (five CTA rows, many CTAs per row) run twice: every state, counter and
    record must be bit-identical (the kernel's cross-CTA traffic -- redundant writes of the accepted station terms,
    partial sums, the grid barrier -- leaves no room for a timing-dependent result), and the carried sums must still
    be the likelihood of the state after ~7000 incremental shared-parameter commits per chain."""
    E, S, R, K, n_it = 3000, 20, 20, 5, 20000
    syn = H.Synthetic(E, S, 12)
    cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=n_it, n_burn=0, n_interval=2000,
                           mode=H.MODE_BLOCKED_GIBBS, precision=32, max_samples=16, step_size_vs=2e-3, step_size_qs=2.0,
                           step_size_t_corr=5e-3, step_size_a_corr=2e-3)
    runs = []
    for _ in range(2):
        with H.HypoTremorB200(cfg) as g:
            g.load(syn)
            g.init_chains()
            g.run(1, 7000)
            g.run(7001, n_it)
            states = [g.get_chain_state(r, k) for r in range(0, R, 3) for k in range(K)]
            L64 = None
            if not runs:
                L64 = g.loglik(np.stack([s["hypo"] for s in states]), np.stack([s["t_corr"] for s in states]),
                               np.stack([s["a_corr"] for s in states]), [s["vs"] for s in states], [s["qs"] for s in states])
            runs.append((states, g.get_counts(), [g.fetch_samples(r) for r in range(R)], L64))
    a, b = runs
    for x, y in zip(a[0], b[0]):
        assert np.array_equal(x["hypo"], y["hypo"]) and x["vs"] == y["vs"] and x["qs"] == y["qs"] and x["temp"] == y["temp"]
        assert np.array_equal(x["t_corr"], y["t_corr"]) and np.array_equal(x["a_corr"], y["a_corr"])
        assert x["log_likelihood"] == y["log_likelihood"]
    assert np.array_equal(a[1][0], b[1][0]) and np.array_equal(a[1][1], b[1][1])
    for x, y in zip(a[2], b[2]):
        assert np.array_equal(x["iter"], y["iter"]) and np.array_equal(x["hypo"], y["hypo"]) and np.array_equal(x["vs"], y["vs"])
    p, acc = a[1]
    assert p[:4].sum() == n_it * R and acc[:4].sum() > 0.05 * p[:4].sum()     # shared parameters do get accepted
    for s, L in zip(a[0], a[3]):
        assert abs(s["log_likelihood"] - L) <= 2e-5 * abs(L), (s["log_likelihood"], L)


def test_blocked_gibbs_float32_result_does_not_depend_on_the_cta_rows(monkeypatch):
    """The launcher deals the chain quads to CTA rows for resident warps per SM (100 chains: 5 rows of 5 warps); the
    shape only regroups which CTA visits which (chain, event), so every decision must be the same with the fewest
    rows (HTM_GIBBS_ROWS=0: 4 rows of 7 + 6 + 6 + 6 warps) and with one quad per row (25 rows); the carried sums are
    float64 sums of the same float32 terms in another grouping."""
    E, S, R, K, n_it = 300, 21, 20, 5, 30
    syn = H.Synthetic(E, S, 9)
    cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=n_it, n_burn=0, n_interval=5,
                           mode=H.MODE_BLOCKED_GIBBS, precision=32, max_samples=8)
    runs = []
    for rows in (None, "0", "25"):
        if rows is None:
            monkeypatch.delenv("HTM_GIBBS_ROWS", raising=False)
        else:
            monkeypatch.setenv("HTM_GIBBS_ROWS", rows)
        with H.HypoTremorB200(cfg) as g:
            g.load(syn)
            g.init_chains()
            tr, sw = g.run_traced(1, n_it)
            runs.append((tr, sw, [g.get_chain_state(r, k) for r in range(R) for k in range(K)], g.get_counts()))
    ref = runs[0]
    assert ref[0]["accepted"].sum() > 0
    for other in runs[1:]:
        for f in ("proposal_type", "index", "prior_ok", "accepted"):
            assert np.array_equal(ref[0][f], other[0][f]), f
        assert np.array_equal(ref[1], other[1])
        assert np.array_equal(ref[3][0], other[3][0]) and np.array_equal(ref[3][1], other[3][1])
        for x, y in zip(ref[2], other[2]):
            assert np.array_equal(x["hypo"], y["hypo"]) and x["vs"] == y["vs"] and x["qs"] == y["qs"] and x["temp"] == y["temp"]
            assert np.array_equal(x["t_corr"], y["t_corr"]) and np.array_equal(x["a_corr"], y["a_corr"])
            assert abs(x["log_likelihood"] - y["log_likelihood"]) <= 1e-12 * abs(x["log_likelihood"])


def test_blocked_gibbs_chunked_runs_equal_one_run():
    syn = H.Synthetic(100, 20, 3)
    cfg = H.default_config(n_sta=20, n_events=100, n_procs=2, n_chains=4, n_iter=80, n_burn=0, n_interval=10,
                           mode=H.MODE_BLOCKED_GIBBS, precision=32)
    with H.HypoTremorB200(cfg) as a, H.HypoTremorB200(cfg) as b:
        for g in (a, b):
            g.load(syn)
            g.init_chains()
        a.run(1, 80)
        b.run(1, 33)
        b.run(34, 80)
        for r in range(2):
            for k in range(4):
                x, y = a.get_chain_state(r, k), b.get_chain_state(r, k)
                assert np.array_equal(x["hypo"], y["hypo"]) and x["vs"] == y["vs"] and x["temp"] == y["temp"]
                assert np.array_equal(x["t_corr"], y["t_corr"]) and x["log_likelihood"] == y["log_likelihood"]
        assert np.array_equal(a.get_counts()[0], b.get_counts()[0])


def test_blocked_gibbs_float32_posterior_matches_reference_schedule_oracle():
    # correctness part (2) with the shared parameters solved (sample-file setting solve_* = T): marginals
    # of x, y, depth, vs, qs, t_corr, a_corr from the B200 float32 blocked-Gibbs kernels vs the oracle's
    # reference schedule (mode A).  Pass: KS p > 1e-3 on thinned samples, medians within 0.2 sigma.
    # (station-term step sizes are raised from the sample file's 0.03 / 0.005 so that these barely
    # constrained parameters decorrelate within the run; both sides use the same settings)
    syn = H.Synthetic(3, 8, 31)
    base = dict(n_sta=8, n_events=3, n_procs=2, n_chains=4, n_cool=1, step_size_t_corr=0.35, step_size_a_corr=0.015)
    cfgA = H.default_config(mode=H.MODE_REPLAY, precision=64, n_iter=6000000, n_interval=97, n_burn=300000, **base)
    o = Oracle(cfgA, syn)
    o.init_chains()
    o.run(1, cfgA.n_iter, trace=False)
    sa = [o.fetch_samples(r) for r in range(2)]
    A = {k: np.concatenate([x[k] for x in sa]) for k in ("vs", "qs", "hypo", "t_corr", "a_corr")}
    n_it, interval, burn = 300000, 7, 20000
    cfgC = H.default_config(mode=H.MODE_BLOCKED_GIBBS, precision=32, n_iter=n_it, n_interval=interval, n_burn=burn,
                            max_samples=4096, **base)
    parts = []
    with H.HypoTremorB200(cfgC) as g:
        g.load(syn)
        g.init_chains()
        it0 = 1
        while it0 <= n_it:                     # drained in chunks, like the Fortran driver does
            it1 = min(n_it, it0 + 4000 * interval - 1)
            g.run(it0, it1)
            for r in range(2):
                parts.append(g.fetch_samples(r))
                g.fetch_likelihood(r)
            it0 = it1 + 1
    C = {k: np.concatenate([x[k] for x in parts]) for k in ("vs", "qs", "hypo", "t_corr", "a_corr")}
    assert len(C["vs"]) == 2 * ((n_it - burn) // interval)
    pick = [("vs", lambda d: d["vs"]), ("qs", lambda d: d["qs"]), ("z0", lambda d: d["hypo"][:, 2]),
            ("x1", lambda d: d["hypo"][:, 3]), ("y2", lambda d: d["hypo"][:, 7]), ("tc0", lambda d: d["t_corr"][:, 0]),
            ("ac3", lambda d: d["a_corr"][:, 3])]
    for name, f in pick:
        a, c = f(A), f(C)
        assert stats.ks_2samp(a[::150], c[::150]).pvalue > 1e-3, name
        assert abs(np.median(a) - np.median(c)) < 0.2 * np.std(a), name


def test_blocked_gibbs_rank_shards_are_independent_ensembles():
    # mode C shards the virtual ranks (the reference's own decomposition): shard 1 of 2 reproduces the oracle's
    # statement with the same global Philox ids, and the two shards draw different streams
    E, S, R, K = 6, 9, 4, 3
    syn = H.Synthetic(E, S, 61)
    base = dict(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=40, n_burn=0, n_interval=5,
                mode=H.MODE_BLOCKED_GIBBS, precision=64, shard_count=2)
    finals = []
    for shard in (0, 1):
        cfg = H.default_config(shard_rank=shard, **base)
        ocfg = H.copy_config(cfg, n_procs=R // 2, shard_count=1, shard_rank=0)
        o = Oracle(ocfg, syn)
        o.set_rank_shard(shard * (R // 2) * K, R * K, shard)
        o.init_chains()
        tr_o, sw_o = o.run(1, 40)
        with H.HypoTremorB200(cfg) as g:
            assert g.n_procs == R // 2 and g.n_events == E
            g.load(syn)
            g.init_chains()
            tr_g, sw_g = g.run_traced(1, 40)
            finals.append(g.get_chain_state(0, 0)["hypo"])
        for f in FLAGS:
            assert np.array_equal(tr_o[f], tr_g[f]), f
        assert np.array_equal(sw_o, sw_g)
        assert rel(tr_g["log_likelihood"], tr_o["log_likelihood"]) <= 1e-9
    assert not np.allclose(finals[0], finals[1])


def test_argument_and_state_errors():
    syn = H.Synthetic(4, 6, 2)
    cfg = fact_cfg(4, 6, 2, 3, max_samples=4, n_interval=10)
    with H.HypoTremorB200(cfg) as g:
        with pytest.raises(H.HtmError) as ei:          # nothing set yet
            g.init_chains()
        assert ei.value.code == H.config.HTM_ERR_STATE
        g.load(syn)
        with pytest.raises(H.HtmError) as ei:          # chains not initialised
            g.run(1, 10)
        assert ei.value.code == H.config.HTM_ERR_STATE
        g.init_chains()
        with pytest.raises(H.HtmError) as ei:
            g.run(0, 10)
        assert ei.value.code == H.config.HTM_ERR_ARG
        g.run(1, 40)                                    # 4 recorded iterations fill the ring
        with pytest.raises(H.HtmError) as ei:          # ring full until fetched
            g.run(41, 80)
        assert ei.value.code == H.config.HTM_ERR_STATE and "sample ring full" in str(ei.value)
        for r in range(2):
            g.fetch_samples(r)
            g.fetch_likelihood(r)
        g.run(41, 80)
        with pytest.raises(H.HtmError):                 # wrong mode
            g.replay(1, 2, [np.zeros(4, dtype=np.int32)] * 2)
    big = fact_cfg(2, 6, 1, 1025)                        # one CTA per tempering group: n_chains <= 1024
    with H.HypoTremorB200(big) as g:
        g.load(H.Synthetic(2, 6, 1))
        g.init_chains()
        with pytest.raises(H.HtmError) as ei:
            g.run(1, 2)
        assert ei.value.code == H.config.HTM_ERR_UNSUPPORTED
    wide = H.default_config(n_sta=400, n_events=2, mode=H.MODE_BLOCKED_GIBBS, precision=32, n_procs=1, n_chains=2)
    with H.HypoTremorB200(wide) as g:                    # blocked-Gibbs stages 32 rows per CTA: n_sta limit
        g.load(H.Synthetic(2, 400, 1))
        with pytest.raises(H.HtmError) as ei:            # (float32: the set-up pass uses the same kernel)
            g.init_chains()
            g.run(1, 2)
        assert ei.value.code == H.config.HTM_ERR_UNSUPPORTED


def test_many_stations_lane_kernel():
    # 1000 stations: the lane kernel falls back to one warp per CTA to fit its shared-memory slice
    syn = H.Synthetic(3, 1000, 4)
    cfg = fact_cfg(3, 1000, 1, 4, precision=64, n_iter=6, n_interval=2)
    o = Oracle(cfg, syn)
    o.init_chains()
    tr_o, _ = o.run(1, 6)
    with H.HypoTremorB200(cfg) as g:
        g.load(syn)
        g.init_chains()
        tr_g, _ = g.run_traced(1, 6)
    for f in FLAGS:
        assert np.array_equal(tr_o[f], tr_g[f])
    assert rel(tr_g["log_likelihood"], tr_o["log_likelihood"]) <= 1e-9
    cfg32 = H.copy_config(cfg, precision=32)
    with H.HypoTremorB200(cfg32) as g:
        g.load(syn)
        g.init_chains()
        g.run(1, 50)
        p, a = g.get_counts()
    assert p[4:].sum() == 50 * 3


def test_abi_level_nccl_gather_single_shard():
    # htm_comm_unique_id / htm_comm_init / htm_gather with NCCL loaded by the library itself (world of one here;
    # tools/comm_check.py runs the same calls on several GPUs under torchrun)
    syn = H.Synthetic(50, 12, 3)
    cfg = fact_cfg(50, 12, 2, 4, n_iter=100, n_interval=10, hist_bins=16)
    with H.HypoTremorB200(cfg) as g:
        g.load(syn)
        g.init_chains()
        g.run(1, 100)
        g.comm_init(H.HypoTremorB200.comm_unique_id())
        hist, p, a = g.gather()
        assert np.array_equal(hist, g.get_histograms())
        p0, a0 = g.get_counts()
        assert np.array_equal(p, p0) and np.array_equal(a, a0)


def test_event_sharded_factorised_gathers_on_two_gpus():
    """Mode B, events sharded over two processes / GPUs: htm_gather (histograms, counters) and htm_gather_samples
    (records with the hypocentres of all events) equal the unsharded run."""
    import subprocess, sys, os, socket
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), "tools/comm_check.py"],
                       cwd=root, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("gathered samples == unsharded samples: True") == 2 and "gathered == unsharded: True" in r.stdout


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_event_sharded_gibbs_on_two_gpus(exchange):
    """Event-sharded joint chains, one process per GPU: the per-iteration exchange of the per-chain sums (fused
    into the sweep over NVLink peer memory, or NCCL) must reproduce the UNSHARDED oracle step by step."""
    import subprocess, sys, os, socket
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, HTM_GIBBS_EXCHANGE=exchange)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), "tests/checks/comm_check_gibbs.py"],
                       cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("matches the unsharded oracle: True") == 2, r.stdout[-2000:]
    if exchange == "nccl":
        return  # float32 event shards exchange through peer memory only
    # float32 (persistent kernel with the exchange between two grid barriers) against the unsharded run
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), "tests/checks/comm_check_gibbs_f32.py"],
                       cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("equals the unsharded run: True") == 2, r.stdout[-2000:]
    # a shard that starts 12 s late: the exchange waits (wall-clock budget, not a poll count) and nothing changes
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), "tests/checks/comm_check_gibbs_f32.py"],
                       cwd=root, env=dict(env, HTM_TEST_DELAY_S="12"), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.count("equals the unsharded run: True") == 2, r.stdout[-2000:] + r.stderr[-2000:]
    # a shard later than the budget: every result-returning call fails loudly on every rank
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), "tests/checks/comm_check_gibbs_f32.py"],
                       cwd=root, env=dict(env, HTM_TEST_DELAY_S="8", HTM_XCH_TIMEOUT_S="2", HTM_TEST_EXPECT_TIMEOUT="1"),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.count("reports the exchange time-out: True") == 2, r.stdout[-2000:] + r.stderr[-2000:]


def test_blocked_gibbs_allreduce_path_is_step_exact(monkeypatch):
    # the event-sharded joint-chain path (sweep -> local totals -> ncclAllReduce -> replicated decide), forced
    # here on a one-rank communicator; tests/checks/comm_check_gibbs.py runs it on several GPUs
    monkeypatch.setenv("HTM_GIBBS_FORCE_ALLREDUCE", "1")
    E, S, R, K = 45, 14, 2, 3
    syn = H.Synthetic(E, S, 71)
    cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=50, n_burn=10, n_interval=5,
                           mode=H.MODE_BLOCKED_GIBBS, precision=64, max_samples=16)
    o = Oracle(cfg, syn)
    o.init_chains()
    tr_o, sw_o = o.run(1, 50)
    with H.HypoTremorB200(cfg) as g:
        g.load(syn)
        g.init_chains()
        g.comm_init(H.HypoTremorB200.comm_unique_id())
        tr_g, sw_g = g.run_traced(1, 50)
        _, nl, _ = g.last_run_stats()
        cg = g.get_counts()
        smp = g.fetch_samples(0)
    assert nl == 1 + 3 * 50
    for f in FLAGS:
        assert np.array_equal(tr_o[f], tr_g[f]), f
    assert np.array_equal(sw_o, sw_g) and rel(tr_g["log_likelihood"], tr_o["log_likelihood"]) <= 1e-9
    co = o.get_counts()
    assert np.array_equal(co[0], cg[0]) and np.array_equal(co[1], cg[1])
    so = o.fetch_samples(0)
    assert np.array_equal(so["iter"], smp["iter"]) and np.allclose(so["vs"], smp["vs"], rtol=1e-12)


@pytest.mark.parametrize("use_time,use_amp", [(1, 0), (0, 1)])
@pytest.mark.parametrize("mode", [H.MODE_FACTORISED, H.MODE_BLOCKED_GIBBS])
def test_data_switches_and_degenerate_sigma_inside_the_kernels(mode, use_time, use_amp):
    # use_time / use_amp drop their sum (src/cls_forward.f90:279,290) and a station with t_stdv = 0 follows the
    # degenerate-sigma rule (:78-90) in the sampling kernels too, not only in htm_loglik
    E, S, R, K = 5, 9, 2, 3
    syn = H.Synthetic(E, S, 88)
    syn.t_stdv[1, 4] = 0.0
    syn.t_stdv[3, 0] = 1e-17
    kw = NOSOLVE if mode == H.MODE_FACTORISED else {}
    cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=50, n_burn=0, n_interval=5, mode=mode,
                           precision=64, use_time=use_time, use_amp=use_amp, max_samples=16, **kw)
    o = Oracle(cfg, syn)
    o.init_chains()
    tr_o, sw_o = o.run(1, 50)
    with H.HypoTremorB200(cfg) as g:
        g.load(syn)
        g.init_chains()
        tr_g, sw_g = g.run_traced(1, 50)
    for f in FLAGS:
        assert np.array_equal(tr_o[f], tr_g[f]), f
    assert np.array_equal(sw_o, sw_g) and rel(tr_g["log_likelihood"], tr_o["log_likelihood"]) <= 1e-9
    # float32 kernels: same configuration runs and its carried likelihood matches a float64 evaluation
    c32 = H.copy_config(cfg, precision=32)
    with H.HypoTremorB200(c32) as g:
        g.load(syn)
        g.init_chains()
        g.run(1, 50)
        st = g.get_chain_state(1, 2)
    L64 = o.loglik(st["hypo"][None, :], st["t_corr"][None, :], st["a_corr"][None, :], [st["vs"]], [st["qs"]])[0]
    assert abs(L64 - st["log_likelihood"]) <= 2e-3 * E + 5e-5 * abs(L64)
