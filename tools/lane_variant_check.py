"""Digest of a factorised-mode float32 run (chain states, counters, thinned samples) for one library build
(HTM_B200_LIB): variants of the lane kernel that only reschedule work must print the same digests as the default."""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import hypotremormcmc_b200 as H

out = {}
for (E, S, R, K, n_it, slots) in ((40, 50, 4, 16, 300, 0), (40, 50, 4, 16, 300, 1), (7, 21, 3, 5, 257, 0), (5, 9, 2, 8, 130, 0),
                                  (3, 1, 2, 4, 64, 0), (6, 40, 4, 16, 100, 2), (6, 44, 2, 32, 100, 2)):
    syn = H.Synthetic(E, S, 5)
    cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=n_it, n_burn=0, n_interval=10,
                           mode=H.MODE_FACTORISED, solve_vs=0, solve_t_corr=0, solve_qs=0, solve_a_corr=0, precision=32,
                           kernel=2, lane_slots=slots, hist_bins=16, max_samples=64)
    h = hashlib.sha256()
    with H.HypoTremorB200(cfg) as g:
        g.load(syn)
        g.init_chains()
        g.run(1, n_it // 2)           # two launches: the words of the first iteration of a launch are not woven
        g.run(n_it // 2 + 1, n_it)
        p, a = g.get_counts()
        h.update(np.asarray(p).tobytes() + np.asarray(a).tobytes())
        for r in range(R):
            s = g.fetch_samples(r)
            for key in sorted(s):
                h.update(np.ascontiguousarray(s[key]).tobytes())
            for j in range(K):
                st = g.get_chain_state(r, j)
                for key in sorted(st):
                    h.update(np.ascontiguousarray(st[key]).tobytes())
    out["E%d_S%d_R%d_K%d_s%d" % (E, S, R, K, slots)] = h.hexdigest()[:16]
print(os.environ.get("HTM_B200_LIB", "default"), json.dumps(out))
