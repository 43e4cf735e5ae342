"""Time the float32 mode C per-iteration sweep in its two layouts (HTM_GIBBS_SWEEP=chain|octet) and the auto rule."""
import json
import os
import sys

sys.path.insert(0, ".")
import hypotremormcmc_b200 as H

os.environ["HTM_GIBBS_PERSIST"] = "0"
out = {}
for (E, S, R, K, n_it) in ((100000, 50, 4, 5, 30), (30000, 50, 4, 5, 60), (10000, 20, 4, 16, 100), (100000, 20, 4, 16, 20),
                           (3000, 50, 4, 5, 200)):
    syn = H.Synthetic(E, S, 5)
    cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=1000, n_burn=0, n_interval=50,
                           mode=H.MODE_BLOCKED_GIBBS, precision=32)
    for cb in os.environ.get("LAYOUTS", "chain,octet,auto").split(","):
        if cb == "auto":
            os.environ.pop("HTM_GIBBS_SWEEP", None)
        else:
            os.environ["HTM_GIBBS_SWEEP"] = cb
        with H.HypoTremorB200(cfg) as g:
            g.load(syn)
            g.init_chains()
            g.run(1, 10)
            g.synchronize()
            best = 1e30
            for rep in range(2):
                g.run(11 + rep * n_it, 10 + (rep + 1) * n_it)
                ms, nl, npr = g.last_run_stats()
                best = min(best, ms)
        out["E%d_S%d_J%d_%s" % (E, S, R * K, cb)] = [round(best * 1e3 / n_it, 1), round(npr / (best * 1e-3) / 1e9, 2)]
print(json.dumps(out))
