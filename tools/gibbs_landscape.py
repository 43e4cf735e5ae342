"""Mode C (blocked Gibbs, float32) time per iteration over event counts, with the library's own path choice."""
import json
import sys

sys.path.insert(0, ".")
import hypotremormcmc_b200 as H

out = {}
for (S, R, K) in ((20, 4, 5), (50, 4, 5), (20, 4, 16)):
    for E in (300, 1000, 3000, 6000, 9000, 12000, 20000, 50000):
        n_it = max(20, min(2000, 4000000 // (E * R * K // 20)))
        syn = H.Synthetic(E, S, 5)
        cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=100000, n_burn=0, n_interval=50,
                               mode=H.MODE_BLOCKED_GIBBS, precision=32)
        with H.HypoTremorB200(cfg) as g:
            g.load(syn)
            g.init_chains()
            g.run(1, 10)
            g.synchronize()
            best, nl = 1e30, 0
            for rep in range(2):
                g.run(11 + rep * n_it, 10 + (rep + 1) * n_it)
                ms, nl, npr = g.last_run_stats()
                best = min(best, ms)
        out["S%d_J%d_E%d" % (S, R * K, E)] = [round(best * 1e3 / n_it, 1), round(npr / (best * 1e-3) / 1e9, 2),
                                               "persistent" if nl <= 2 else "per-iteration"]
        print("S%d_J%d_E%d" % (S, R * K, E), out["S%d_J%d_E%d" % (S, R * K, E)], flush=True)
