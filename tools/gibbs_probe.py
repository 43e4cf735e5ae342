import sys
sys.path.insert(0, ".")
import numpy as np
import hypotremormcmc_b200 as H
E, S, R, K = int(sys.argv[1]), int(sys.argv[2]), 4, 5
n_it = int(sys.argv[3])
syn = H.Synthetic(E, S, 5)
cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=n_it, n_burn=0, n_interval=50,
                       mode=H.MODE_BLOCKED_GIBBS, precision=32)
with H.HypoTremorB200(cfg) as g:
    g.load(syn)
    g.init_chains()
    g.run(1, 20)
    g.synchronize()
    g.run(21, 20 + n_it)
    ms, nl, npr = g.last_run_stats()
    print("E=%d S=%d: %.2f us/iter, %.3g proposals/s" % (E, S, ms * 1e3 / n_it, npr / (ms * 1e-3)))
