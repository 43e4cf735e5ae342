"""Small run through every kernel (for compute-sanitizer)."""
import sys
sys.path.insert(0, ".")
import os
import numpy as np
import hypotremormcmc_b200 as H
from oracle.pyoracle import Oracle

NOSOLVE = dict(solve_vs=0, solve_t_corr=0, solve_qs=0, solve_a_corr=0)
syn = H.Synthetic(5, 11, 3)
for prec in (32, 64):
    for kernel, slots in ((2, 1), (2, 2), (1, 0)):
        cfg = H.default_config(n_sta=11, n_events=5, n_procs=3, n_chains=5, n_cool=1, n_iter=30, n_burn=0, n_interval=5,
                               mode=H.MODE_FACTORISED, precision=prec, kernel=kernel, lane_slots=slots, max_samples=8,
                               hist_bins=8, **NOSOLVE)
        with H.HypoTremorB200(cfg) as g:
            g.load(syn)
            g.init_chains()
            g.run(1, 15)
            g.run_traced(16, 30)
            g.fetch_samples(0), g.get_histograms(), g.get_counts()
    for persist in ("0", "1"):
        os.environ["HTM_GIBBS_PERSIST"] = persist
        cfg = H.default_config(n_sta=11, n_events=5, n_procs=2, n_chains=5, n_cool=1, n_iter=30, n_burn=0, n_interval=5,
                               mode=H.MODE_BLOCKED_GIBBS, precision=prec, max_samples=8)
        with H.HypoTremorB200(cfg) as g:
            g.load(syn)
            g.init_chains()
            g.run(1, 15)
            g.run_traced(16, 30)
            g.fetch_samples(1), g.fetch_likelihood(0)
    cfg = H.default_config(n_sta=11, n_events=5, mode=H.MODE_REPLAY if prec == 64 else H.MODE_FACTORISED, precision=prec,
                           **(dict() if prec == 64 else NOSOLVE))
    with H.HypoTremorB200(cfg) as g:
        g.load(syn)
        rng = np.random.default_rng(0)
        g.loglik(rng.normal(0, 10, (3, 15)) + 8, np.zeros((3, 11)), np.zeros((3, 11)), np.full(3, 3.0), np.full(3, 250.0))
cfgA = H.default_config(n_sta=11, n_events=5, n_procs=2, n_chains=3, n_iter=40, n_burn=0, n_interval=10, mode=H.MODE_REPLAY,
                        precision=64)
o = Oracle(cfgA, syn)
o.init_chains()
st = [[o.get_chain_state(r, j) for j in range(3)] for r in range(2)]
o.record_draws(True)
o.run(1, 40)
with H.HypoTremorB200(cfgA) as g:
    g.load(syn)
    for r in range(2):
        for j in range(3):
            s = st[r][j]
            g.set_chain_state(r, j, s["hypo"], s["t_corr"], s["a_corr"], s["vs"], s["qs"], s["temp"], s["log_likelihood"])
    g.replay(1, 40, [o.draws(0), o.draws(1)])
print("sanitize probe ok")
