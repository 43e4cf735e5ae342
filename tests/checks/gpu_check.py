"""Quick GPU sanity + perf sweep (development aid; the formal checks are tests/ and bench.py)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import hypotremormcmc_b200 as H
from hypotremormcmc_b200.api import measure_fp32_peak
from oracle.pyoracle import Oracle

out = {}


def fact_cfg(E, S, R, K, **kw):
    return H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=1000, n_burn=0,
                            n_interval=10, mode=H.MODE_FACTORISED, solve_vs=0, solve_t_corr=0, solve_qs=0,
                            solve_a_corr=0, **kw)


def check_loglik():
    syn = H.Synthetic(50, 20, 7)
    rng = np.random.default_rng(1)
    M = 8
    hypo = np.stack([np.stack([syn.true_x + rng.normal(0, 3, 50), syn.true_y + rng.normal(0, 3, 50),
                               syn.true_z + rng.normal(0, 1, 50)], axis=1).ravel() for _ in range(M)])
    tc = rng.normal(0, 0.1, (M, 20)); ac = rng.normal(0, 0.01, (M, 20))
    vs = rng.uniform(2.5, 3.5, M); qs = rng.uniform(150, 350, M)
    for prec in (64, 32):
        cfg = H.default_config(n_sta=20, n_events=50, mode=H.MODE_REPLAY if prec == 64 else H.MODE_FACTORISED,
                               precision=prec, solve_vs=0, solve_t_corr=0, solve_qs=0, solve_a_corr=0)
        o = Oracle(cfg, syn)
        Lo, peo = o.loglik(hypo, tc, ac, vs, qs, per_event=True)
        with H.HypoTremorB200(cfg) as g:
            g.load(syn)
            Lg, peg = g.loglik(hypo, tc, ac, vs, qs, per_event=True)
        rel = np.max(np.abs(peg - peo) / np.maximum(1, np.abs(peo)))
        out["loglik_f%d_max_rel_per_event" % prec] = float(rel)
        out["loglik_f%d_max_abs_per_event" % prec] = float(np.max(np.abs(peg - peo)))
        out["loglik_f%d_total_rel" % prec] = float(np.max(np.abs(Lg - Lo) / np.abs(Lo)))


def check_replay():
    syn = H.Synthetic(10, 12, 11)
    cfg = H.default_config(n_sta=12, n_events=10, n_procs=3, n_chains=4, n_iter=400, n_burn=100, n_interval=10,
                           mode=H.MODE_REPLAY, precision=64)
    o = Oracle(cfg, syn)
    o.init_chains()
    states = [[o.get_chain_state(r, j) for j in range(4)] for r in range(3)]
    o.record_draws(True)
    tr_o, sw_o = o.run(1, 400)
    draws = [o.draws(r) for r in range(3)]
    with H.HypoTremorB200(cfg) as g:
        g.load(syn)
        for r in range(3):
            for j in range(4):
                s = states[r][j]
                g.set_chain_state(r, j, s["hypo"], s["t_corr"], s["a_corr"], s["vs"], s["qs"], s["temp"],
                                  s["log_likelihood"])
        tr_g, sw_g, used = g.replay(1, 400, draws)
        cg = g.get_counts()
    same = all(np.array_equal(tr_o[f], tr_g[f]) for f in ("proposal_type", "index", "prior_ok", "accepted"))
    dl = np.abs(tr_o["log_likelihood"] - tr_g["log_likelihood"]) / np.maximum(1, np.abs(tr_o["log_likelihood"]))
    out["replay_flags_equal"] = bool(same)
    out["replay_L_max_rel"] = float(dl.max())
    out["replay_swaps_equal"] = bool(np.array_equal(sw_o, sw_g))
    out["replay_draws_used_equal"] = bool(np.array_equal(used, [len(d) for d in draws]))
    co = o.get_counts()
    out["replay_counts_equal"] = bool(np.array_equal(co[0], cg[0]) and np.array_equal(co[1], cg[1]))


def check_factorised():
    for S in (10, 50):
        syn = H.Synthetic(6, S, 13)
        for kernel, slots in ((2, 1), (2, 2), (2, 4), (1, 0)):
            cfg = fact_cfg(6, S, 4, 5, precision=64, kernel=kernel, lane_slots=slots, max_samples=16, hist_bins=32)
            o = Oracle(cfg, syn)
            o.init_chains()
            so = o.factorised_state()
            tr_o, sw_o = o.run(1, 60)
            with H.HypoTremorB200(cfg) as g:
                g.load(syn)
                g.init_chains()
                st0 = g.get_chain_state(0, 0)
                tr_g, sw_g = g.run_traced(1, 60)
                cg = g.get_counts()
                smp = g.fetch_samples(1)
                lk = g.fetch_likelihood(1)
            key = "fact_S%d_k%d_s%d" % (S, kernel, slots)
            out[key + "_init_x_err"] = float(np.max(np.abs(st0["hypo"][0::3] - so["x"][:, 0, 0])))
            flags = all(np.array_equal(tr_o[f], tr_g[f]) for f in ("proposal_type", "index", "prior_ok", "accepted"))
            out[key + "_flags_equal"] = bool(flags)
            if not flags:
                out[key + "_n_flag_diff"] = int(np.sum(tr_o["accepted"] != tr_g["accepted"]))
            out[key + "_L_max_rel"] = float(np.max(np.abs(tr_o["log_likelihood"] - tr_g["log_likelihood"]) /
                                                   np.maximum(1, np.abs(tr_o["log_likelihood"]))))
            out[key + "_swaps_equal"] = bool(np.array_equal(sw_o, sw_g))
            co = o.get_counts()
            out[key + "_counts_equal"] = bool(np.array_equal(co[0], cg[0]) and np.array_equal(co[1], cg[1]))
            so_s = o.fetch_samples(1)
            out[key + "_samples_equal"] = bool(np.array_equal(so_s["iter"], smp["iter"]) and
                                               np.allclose(so_s["hypo"], smp["hypo"], rtol=1e-9, atol=1e-9))
            lo = o.fetch_likelihood(1)
            out[key + "_lik_equal"] = bool(np.array_equal(lo[0], lk[0]) and np.allclose(lo[1], lk[1], rtol=1e-9))


def perf():
    for (E, S, R, K, n_it) in ((1000, 20, 4, 16, 2000), (10000, 50, 4, 16, 500), (1000, 10, 4, 5, 2000)):
        syn = H.Synthetic(E, S, 5)
        for kernel, slots in ((2, 1), (2, 2), (2, 4), (1, 0)):
            if kernel == 1 and K > 16:
                continue
            cfg = fact_cfg(E, S, R, K, precision=32, kernel=kernel, lane_slots=slots, hist_bins=64,
                           ladder=H.LADDER_GEOMETRIC)
            cfg.n_interval = 100
            try:
                with H.HypoTremorB200(cfg) as g:
                    g.load(syn)
                    g.init_chains()
                    g.run(1, 200)
                    g.synchronize()
                    best = 1e30
                    for rep in range(3):
                        g.run(201 + rep * n_it, 200 + (rep + 1) * n_it)
                        ms, nl, npr = g.last_run_stats()
                        best = min(best, ms)
                    p, a = g.get_counts()
                out["perf_E%d_S%d_R%d_K%d_k%d_s%d" % (E, S, R, K, kernel, slots)] = dict(
                    ms=best, proposals_per_s=npr / (best * 1e-3), accept_rate=float(a.sum() / max(1, p.sum())))
            except H.HtmError as ex:
                out["perf_E%d_S%d_k%d_s%d" % (E, S, kernel, slots)] = str(ex)


def perf_gibbs():
    for (E, S, R, K, n_it) in ((1000, 20, 4, 5, 300), (10000, 50, 4, 5, 100), (100000, 50, 4, 5, 20)):
        syn = H.Synthetic(E, S, 5)
        cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=n_it, n_burn=0,
                               n_interval=50, mode=H.MODE_BLOCKED_GIBBS, precision=32)
        with H.HypoTremorB200(cfg) as g:
            g.load(syn)
            g.init_chains()
            g.run(1, 20)
            g.synchronize()
            best = 1e30
            for rep in range(2):
                g.run(21 + rep * n_it, 20 + (rep + 1) * n_it)
                ms, nl, npr = g.last_run_stats()
                best = min(best, ms)
            p, a = g.get_counts()
        out["perf_gibbs_E%d_S%d_J%d" % (E, S, R * K)] = dict(ms=best, us_per_iter=best * 1e3 / n_it,
                                                             proposals_per_s=npr / (best * 1e-3),
                                                             accept=[float(x) for x in (a / np.maximum(1, p))])


if __name__ == "__main__":
    t0 = time.time()
    for fn in (lambda: out.update(fp32_peak=measure_fp32_peak(0)), check_loglik, check_replay, check_factorised, perf, perf_gibbs):
        try:
            fn()
        except Exception as ex:  # keep going: this is a survey
            import traceback
            out["error_" + getattr(fn, "__name__", "peak")] = traceback.format_exc()[-1500:]
    out["seconds"] = time.time() - t0
    print(json.dumps(out, indent=1))
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/gpu_check.json", "w"), indent=1)
