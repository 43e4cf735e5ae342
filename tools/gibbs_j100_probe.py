import sys
sys.path.insert(0, ".")
import hypotremormcmc_b200 as H
for (E, S, R, K) in ((1, 10, 4, 5), (1, 10, 20, 5), (100, 20, 20, 5), (1000, 20, 20, 5), (10000, 20, 20, 5), (10000, 50, 20, 5)):
    n_it = max(40, min(4000, 8000000 // (E * R * K)))
    syn = H.Synthetic(E, S, 5)
    cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=100000, n_burn=0, n_interval=1000,
                           mode=H.MODE_BLOCKED_GIBBS, precision=32)
    with H.HypoTremorB200(cfg) as g:
        g.load(syn); g.init_chains(); g.run(1, 10); g.synchronize()
        best = 1e30
        for rep in range(2):
            g.run(11 + rep * n_it, 10 + (rep + 1) * n_it)
            ms, nl, npr = g.last_run_stats(); best = min(best, ms)
    print("E=%d S=%d J=%d: %.1f us/iter, %.3g proposals/s, launches %d" % (E, S, R * K, best * 1e3 / n_it, npr / (best * 1e-3), nl), flush=True)
