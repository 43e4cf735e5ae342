"""htm_loglik (batched calc_log_likelihood) at a size where it is HBM-bound: time it under
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum` to read the achieved GB/s."""
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import hypotremormcmc_b200 as H

E, S, M = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
prec = int(sys.argv[4]) if len(sys.argv) > 4 else 32
syn = H.Synthetic(E, S, 5)
cfg = H.default_config(n_sta=S, n_events=E, n_procs=1, n_chains=2, n_iter=10, n_interval=5, mode=H.MODE_FACTORISED,
                       solve_vs=0, solve_t_corr=0, solve_qs=0, solve_a_corr=0, precision=prec)
rng = np.random.default_rng(3)
true = np.stack([syn.true_x, syn.true_y, syn.true_z], axis=1).reshape(1, -1)
hypo = np.tile(true, (M, 1)) + rng.normal(0, 0.5, (M, 3 * E))
hypo[:, 2::3] = np.abs(hypo[:, 2::3]) + 1.0
tc, ac = rng.normal(0, 0.05, (M, S)), rng.normal(0, 0.05, (M, S))
vs, qs = np.full(M, 3.0), np.full(M, 250.0)
with H.HypoTremorB200(cfg) as g:
    g.load(syn)
    for _ in range(3):
        t0 = time.perf_counter()
        L = g.loglik(hypo, tc, ac, vs, qs)
        dt = time.perf_counter() - t0
    print("loglik E=%d S=%d M=%d f%d: host call %.2f ms (includes H2D of the models), L[0]=%.6g" % (E, S, M, prec, dt * 1e3, L[0]))
