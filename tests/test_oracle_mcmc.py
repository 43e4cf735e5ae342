"""cls_model / cls_mcmc / cls_parallel / driver-loop restatement: prior identities, draw
schedule, the reference's quirks, output bookkeeping, thread = serial."""
import math

import numpy as np

import hypotremormcmc_b200 as H
from oracle import pyoracle as po
from oracle.pyoracle import Oracle


def small(mode=H.MODE_REPLAY, **kw):
    syn = H.Synthetic(kw.pop("E", 3), kw.pop("S", 8), kw.pop("seed", 9))
    base = dict(n_sta=syn.n_sta, n_events=syn.n_events, n_procs=2, n_chains=3, n_cool=1, n_iter=500, n_burn=100,
                n_interval=10, mode=mode, precision=64)
    base.update(kw)
    return syn, H.default_config(**base)


# ---- cls_model ---------------------------------------------------------------------------------
def test_gaussian_prior_ratio_is_antisymmetric():
    xn, lpr, ok = po.perturb(1.3, 0.5, 2.0, 0.7, 0, 0.9)
    assert ok and xn == 1.3 + 0.9 * 0.7
    # perturbing back (old <-> new) flips the sign of the log prior ratio
    xb, lpr_b, _ = po.perturb(xn, 0.5, 2.0, 0.7, 0, -0.9)
    assert abs(xb - 1.3) < 1e-15 and abs(lpr + lpr_b) < 1e-15
    assert abs(lpr - (-((xn - 0.5) ** 2 - (1.3 - 0.5) ** 2) / (2 * 2.0 * 2.0))) < 1e-16


def test_rayleigh_prior_rejects_below_mu_and_adds_log_terms():
    xn, lpr, ok = po.perturb(0.3, 0.0, 10.0, 0.4, 1, -1.0)  # x_new = -0.1 <= mu
    assert not ok and lpr == float(np.float32(-1.0e30))      # single-precision literal, cls_model.f90:180
    xn, lpr, ok = po.perturb(5.0, 0.0, 10.0, 0.4, 1, 1.0)
    assert ok
    expect = -((5.4) ** 2 - 25.0) / 200.0 + math.log(5.4) - math.log(5.0)
    assert abs(lpr - expect) < 1e-15


def test_swap_with_equal_temperatures_is_always_accepted():
    # del_s = 0 => log(r) <= 0 for every r in [eps, 1)  (cls_parallel.f90:292-299)
    for r in (0.999999, 0.5, 1e-9):
        assert po.judge_swap(3.0, 3.0, -10.0, -500.0, r)
    assert not po.judge_swap(1.0, 50.0, -10.0, -500.0, 0.5)     # moving the worse state to T=1 is penalised
    assert po.judge_swap(1.0, 50.0, -500.0, -10.0, 0.5)
    assert not po.judge_swap(3.0, 3.0, -10.0, -500.0, 1e-17)     # r < eps is never accepted (:295)


# ---- chain set-up (hypo_tremor_mcmc.f90:120-211) ------------------------------------------------
def test_initial_state_and_draw_count():
    syn, cfg = small()
    o = Oracle(cfg, syn)
    o.record_draws(True)
    o.init_chains()
    E, S = syn.n_events, syn.n_sta
    # draws per chain: 2S (t_corr) + 2S (a_corr) + 5E (hypo) + [j > n_cool]
    per_rank = 3 * (2 * S + 2 * S + 5 * E) + 2
    assert len(o.draws(0)) == per_rank and len(o.draws(1)) == per_rank
    for r in range(2):
        for j in range(3):
            st = o.get_chain_state(r, j)
            assert st["vs"] == cfg.prior_vs and st["qs"] == cfg.prior_qs    # start at the prior mean, no draw
            assert st["log_likelihood"] == -9.0e300                          # cls_mcmc.f90:88
            assert np.all(st["hypo"][2::3] > cfg.prior_z)                    # Rayleigh start: z > prior_z
            if j < cfg.n_cool:
                assert st["temp"] == 1.0
            else:
                assert 1.0 <= st["temp"] <= cfg.temp_high


def test_unsolved_station_terms_stay_at_prior_mean_and_cost_no_draws():
    syn, cfg = small(solve_t_corr=0, solve_a_corr=0, prior_t_corr=0.25)
    o = Oracle(cfg, syn)
    o.record_draws(True)
    o.init_chains()
    assert len(o.draws(0)) == 3 * 5 * syn.n_events + 2
    assert np.all(o.get_chain_state(0, 0)["t_corr"] == 0.25)


# ---- one iteration (hypo_tremor_mcmc.f90:236-284) ------------------------------------------------
def test_first_proposal_is_accepted_through_the_sentinel():
    # quirk Q3: L starts at -9e300, so the first prior_ok proposal is always accepted
    syn, cfg = small()
    o = Oracle(cfg, syn)
    o.init_chains()
    tr, sw = o.run(1, 1)
    ok = tr["prior_ok"][0] == 1
    assert np.all(tr["accepted"][0][ok] == 1)
    assert np.all(tr["log_likelihood"][0][ok] > -1e300)


def test_proposal_type_index_mapping_and_counters():
    # quirk Q1: icmp 0 -> z (index 3id), 1 -> y, 2 -> x; proposal type 5 + icmp
    syn, cfg = small(n_iter=3000)
    o = Oracle(cfg, syn)
    o.init_chains()
    tr, sw = o.run(1, 3000)
    t, idx = tr["proposal_type"].ravel(), tr["index"].ravel()
    hyp = t >= 5
    icmp = t[hyp] - 5
    assert np.all((3 * ((idx[hyp] + 2) // 3) - idx[hyp]) == icmp)   # index = 3*id - icmp
    assert np.all(idx[t == 1] == 1) and np.all(idx[t == 3] == 1)
    assert np.all((idx[t == 2] >= 1) & (idx[t == 2] <= syn.n_sta))
    # only z moves (type 5) can fail the prior
    assert np.all(t[tr["prior_ok"].ravel() == 0] == 5)
    # proposal mix: 2.5 % each global type, 90 % hypocentre
    frac = np.array([(t == k).mean() for k in range(1, 8)])
    assert np.allclose(frac[:4], 0.025, atol=0.006) and abs(frac[4:].sum() - 0.9) < 0.01
    # counters only count T = 1 chains (quirk Q9): at most one cold chain per rank on average
    p, a = o.get_counts()
    assert p.sum() < 3000 * 6 and np.all(a <= p)


def test_draw_schedule_three_to_six_per_step():
    syn, cfg = small(n_iter=400)
    o = Oracle(cfg, syn)
    o.init_chains()
    o.record_draws(True)
    tr, sw = o.run(1, 400)
    t, ok = tr["proposal_type"], tr["prior_ok"]
    per_step = 1 + np.where(t >= 5, 2, np.where((t == 2) | (t == 4), 1, 0)) + 2 + ok
    assert per_step.min() >= 3 and per_step.max() <= 6
    for r in range(2):
        chain_draws = per_step[:, r, :].sum()
        swap_draws = np.sum(sw["rank1"] == r)              # judge_swap draws on rank1's stream
        extra = len(o.draws(r)) - chain_draws - swap_draws
        if r == 0:
            assert extra >= 2 * 400                        # select_pair: >= 2 draws per iteration on rank 0
        else:
            assert extra == 0


def test_swaps_exchange_temperatures_only():
    syn, cfg = small(n_iter=600)
    o = Oracle(cfg, syn)
    o.init_chains()
    t0 = sorted(o.get_chain_state(r, j)["temp"] for r in range(2) for j in range(3))
    tr, sw = o.run(1, 600)
    t1 = sorted(o.get_chain_state(r, j)["temp"] for r in range(2) for j in range(3))
    assert t0 == t1 and sw["accepted"].sum() > 0
    assert np.all((sw["rank1"] != sw["rank2"]) | (sw["chain1"] != sw["chain2"]))


def test_recording_rules():
    # T=1 chains, mod(i, n_interval) == 1; samples only after burn-in, likelihood always (Q6)
    syn, cfg = small(n_iter=500, n_burn=200, n_interval=10)
    o = Oracle(cfg, syn)
    o.init_chains()
    o.run(1, 500, trace=False)
    n_s = n_l = 0
    for r in range(2):
        s = o.fetch_samples(r)
        it, lik = o.fetch_likelihood(r)
        assert np.all(s["iter"] % 10 == 1) and np.all(s["iter"] > 200)
        assert np.all(it % 10 == 1) and it.min() == 1
        assert s["hypo"].shape[1] == 3 * syn.n_events and s["t_corr"].shape[1] == syn.n_sta
        n_s += len(s["iter"])
        n_l += len(it)
    # total over ranks is fixed (n_procs * n_cool cold chains), per-rank counts are not (Q10)
    assert n_s == 2 * 30 and n_l == 2 * 50


def test_n_interval_one_records_nothing():
    syn, cfg = small(n_iter=50, n_burn=0, n_interval=1)
    o = Oracle(cfg, syn)
    o.init_chains()
    o.run(1, 50, trace=False)
    assert len(o.fetch_samples(0)["iter"]) == 0 and len(o.fetch_likelihood(0)[0]) == 0


def test_threaded_run_equals_serial_run():
    syn, cfg = small(n_iter=800, n_procs=4, n_chains=2)
    a, b = Oracle(cfg, syn), Oracle(cfg, syn)
    a.init_chains()
    b.init_chains()
    a.run(1, 800, trace=False)
    b.run_threaded(1, 800)
    for r in range(4):
        for j in range(2):
            sa, sb = a.get_chain_state(r, j), b.get_chain_state(r, j)
            assert np.array_equal(sa["hypo"], sb["hypo"]) and sa["temp"] == sb["temp"]
            assert sa["log_likelihood"] == sb["log_likelihood"]
    assert np.array_equal(a.get_counts()[0], b.get_counts()[0])


def test_chunked_run_equals_single_run():
    syn, cfg = small(n_iter=300)
    a, b = Oracle(cfg, syn), Oracle(cfg, syn)
    a.init_chains()
    b.init_chains()
    ta, _ = a.run(1, 300)
    t1, _ = b.run(1, 120)
    t2, _ = b.run(121, 300)
    assert np.array_equal(ta, np.concatenate([t1, t2]))


# ---- factorised schedule (the B200 mode-B statement) ---------------------------------------------
def test_factorised_oracle_bookkeeping():
    syn, cfg = small(mode=H.MODE_FACTORISED, solve_vs=0, solve_t_corr=0, solve_qs=0, solve_a_corr=0, n_iter=400,
                     n_burn=100, n_interval=10, n_chains=4)
    o = Oracle(cfg, syn)
    o.init_chains()
    st = o.factorised_state()
    assert np.all(st["T"][:, :, 0] == 1.0) and np.all(st["T"][:, :, 1:] >= 1.0)
    tr, sw = o.run(1, 400)
    p, a = o.get_counts()
    assert p[:4].sum() == 0 and p[4:].sum() == 400 * syn.n_events * 2     # one cold chain per group
    s = o.fetch_samples(1)
    assert len(s["iter"]) == 30 and np.all(s["vs"] == cfg.prior_vs)
    it, lik = o.fetch_likelihood(1)
    assert len(it) == 40
    st1 = o.factorised_state()
    assert np.array_equal(np.sort(st["T"], axis=2), np.sort(st1["T"], axis=2))   # temperatures only move


def test_factorised_sharding_is_invariant():
    # Philox streams are keyed by GLOBAL event ids, so a shard reproduces its slice of the full run
    syn, cfg = small(mode=H.MODE_FACTORISED, solve_vs=0, solve_t_corr=0, solve_qs=0, solve_a_corr=0, n_iter=100,
                     E=5)
    full = Oracle(cfg, syn)
    full.init_chains()
    tf, sf = full.run(1, 100)
    sh = syn.shard(1, 2)
    part = Oracle(cfg, sh, event_offset=sh.event_offset)
    part.init_chains()
    tp, sp = part.run(1, 100)
    lo = sh.event_offset
    for f in ("proposal_type", "prior_ok", "accepted", "log_likelihood"):
        assert np.array_equal(tf[f][:, lo:lo + sh.n_events], tp[f])
    assert np.array_equal(sf[:, lo:lo + sh.n_events], sp)


def test_blocked_gibbs_oracle_bookkeeping():
    """Mode C statement: per iteration every cold chain proposes one hypocentre coordinate PER EVENT and one shared
    parameter; records follow the reference's rules; temperatures only move; chunked == single run."""
    syn, cfg = small(mode=H.MODE_BLOCKED_GIBBS, n_iter=300, n_burn=100, n_interval=10, n_chains=3, E=4)
    R, K, E = cfg.n_procs, cfg.n_chains, syn.n_events
    o = Oracle(cfg, syn)
    o.init_chains()
    temps0 = sorted(o.get_chain_state(r, k)["temp"] for r in range(R) for k in range(K))
    tr, sw = o.run(1, 300)
    p, a = o.get_counts()
    n_cold = R * cfg.n_cool
    assert p[4:].sum() == 300 * E * n_cold and p[:4].sum() == 300 * n_cold and (a <= p).all()
    # trace: [iteration][event 0..E-1, then the shared-parameter row][chain]
    assert tr["proposal_type"].shape == (300, E + 1, R, K)
    assert set(np.unique(tr["proposal_type"][:, :E])) <= {5, 6, 7} and set(np.unique(tr["proposal_type"][:, E])) <= {1, 2, 3, 4}
    smp = [o.fetch_samples(r) for r in range(R)]
    lik = [o.fetch_likelihood(r) for r in range(R)]
    assert sum(len(s["iter"]) for s in smp) == 20 * n_cold        # iterations 101, 111, ..., 291
    assert sum(len(l[0]) for l in lik) == 30 * n_cold             # the likelihood file includes the burn-in
    assert sorted(o.get_chain_state(r, k)["temp"] for r in range(R) for k in range(K)) == temps0
    # the carried total is the likelihood of the state
    for r in range(R):
        for k in range(K):
            st = o.get_chain_state(r, k)
            L = o.loglik(st["hypo"][None], st["t_corr"][None], st["a_corr"][None], [st["vs"]], [st["qs"]])[0]
            assert abs(L - st["log_likelihood"]) <= 1e-9 * max(1.0, abs(L))
    # chunking does not change anything
    o2 = Oracle(cfg, syn)
    o2.init_chains()
    t1, s1 = o2.run(1, 77)
    t2, s2 = o2.run(78, 300)
    for f in ("proposal_type", "index", "prior_ok", "accepted", "log_likelihood"):
        assert np.array_equal(np.concatenate([t1[f], t2[f]]), tr[f]), f
    assert np.array_equal(np.concatenate([s1, s2]), sw)
