"""Time the factorised lane kernel for one library build (HTM_B200_LIB) on a few workloads."""
import json
import os
import sys

sys.path.insert(0, ".")
import hypotremormcmc_b200 as H

out = {}
for (E, S, R, K, n_it) in ((1000, 20, 4, 16, 4000), (10000, 50, 4, 16, 500), (100000, 50, 2, 16, 60)):
    syn = H.Synthetic(E, S, 5)
    for slots in [int(v) for v in os.environ.get("SLOTS", "0").split(",")]:  # 0 = the library's own choice
        cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=n_it, n_burn=0,
                               n_interval=100, mode=H.MODE_FACTORISED, solve_vs=0, solve_t_corr=0, solve_qs=0,
                               solve_a_corr=0, precision=32, kernel=2, lane_slots=slots, hist_bins=32)
        with H.HypoTremorB200(cfg) as g:
            g.load(syn)
            g.init_chains()
            g.run(1, 100)
            g.synchronize()
            best = 1e30
            for rep in range(3):
                g.run(101 + rep * n_it, 100 + (rep + 1) * n_it)
                ms, nl, npr = g.last_run_stats()
                best = min(best, ms)
        out["E%d_S%d_s%d" % (E, S, slots)] = round(npr / (best * 1e-3) / 1e9, 2)
print(os.environ.get("HTM_B200_LIB", "default"), json.dumps(out))
