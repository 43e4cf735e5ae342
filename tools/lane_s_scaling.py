"""Lane kernel (factorised mode, float32): time per iteration against the number of stations, to split one
warp-iteration into the station loop (slope) and everything else (intercept).  HTM_B200_LIB picks the build."""
import json
import os
import sys

sys.path.insert(0, ".")
import hypotremormcmc_b200 as H

out = {}
E, R, K = 10000, 4, 16
for slots in (1, 2):
    for S in (10, 20, 50, 100, 150, 200):
        n_it = max(100, 25000 // S)
        syn = H.Synthetic(E, S, 5)
        cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=4 * n_it + 100, n_burn=0,
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
        rate = npr / (best * 1e-3)
        # cycles of one scheduler per warp-iteration (32 * slots proposals) at 1965 MHz, 592 schedulers
        cyc = 1.965e9 * 592 * 32 * slots / rate
        out["s%d_S%d" % (slots, S)] = [round(rate / 1e9, 2), round(cyc, 1)]
print(os.environ.get("HTM_B200_LIB", "default"), json.dumps(out))
