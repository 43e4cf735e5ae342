"""mode C (float32) time per iteration at one size: gibbs_probe.py E S iters [R K]"""
import sys
sys.path.insert(0, ".")
import hypotremormcmc_b200 as H
E, S, n_it = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
R, K = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (4, 5)
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
    p, a = g.get_counts()
    print("E=%d S=%d J=%d: %.2f us/iter, %.3g proposals/s, algorithmic %.1f TFLOP/s; accept hypo %.3f shared %.3f"
          % (E, S, R * K, ms * 1e3 / n_it, npr / (ms * 1e-3), npr / (ms * 1e-3) * (60 * S + 64) / 1e12,
             a[4:].sum() / max(1, p[4:].sum()), a[:4].sum() / max(1, p[:4].sum())))
