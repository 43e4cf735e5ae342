"""Long factorised and blocked-Gibbs runs in uneven chunks: counters, records and final states must not depend on
the chunking, and every recorded value must be finite."""
import sys
import numpy as np

sys.path.insert(0, ".")
import hypotremormcmc_b200 as H


def run(cfg, syn, chunks):
    with H.HypoTremorB200(cfg) as g:
        g.load(syn)
        g.init_chains()
        it, smp = 0, [[] for _ in range(cfg.n_procs)]
        for n in chunks:
            g.run(it + 1, it + n)
            it += n
            for r in range(cfg.n_procs):
                s = g.fetch_samples(r)
                smp[r].append((s["iter"].copy(), s["hypo"].copy()))
            g.discard_samples()
        st = [g.get_chain_state(r, k) for r in range(cfg.n_procs) for k in range(cfg.n_chains)]
        return g.get_counts(), smp, st


def check(name, cfg, syn, total):
    a = run(cfg, syn, [total])
    rng = np.random.default_rng(1)
    parts = []
    left = total
    while left > 0:
        n = int(min(left, rng.integers(1, total // 6)))
        parts.append(n)
        left -= n
    b = run(cfg, syn, parts)
    assert np.array_equal(a[0][0], b[0][0]) and np.array_equal(a[0][1], b[0][1]), "counters depend on chunking"
    for r in range(cfg.n_procs):
        ia = np.concatenate([x[0] for x in a[1][r]]); ib = np.concatenate([x[0] for x in b[1][r]])
        ha = np.concatenate([x[1] for x in a[1][r]]); hb = np.concatenate([x[1] for x in b[1][r]])
        assert np.array_equal(ia, ib) and np.array_equal(ha, hb), "records depend on chunking"
        assert np.isfinite(ha).all()
    for x, y in zip(a[2], b[2]):
        assert np.array_equal(x["hypo"], y["hypo"]) and x["log_likelihood"] == y["log_likelihood"] and x["temp"] == y["temp"]
    print(name, "ok:", total, "iterations in", len(parts), "chunks; cold proposals", int(a[0][0].sum()),
          "accept rate %.3f" % (a[0][1].sum() / a[0][0].sum()), flush=True)


syn = H.Synthetic(1000, 20, 5)
check("mode B 1000x20x16x4", H.default_config(n_sta=20, n_events=1000, n_procs=4, n_chains=16, n_cool=1, n_iter=300000, n_burn=1000,
      n_interval=1000, mode=H.MODE_FACTORISED, solve_vs=0, solve_t_corr=0, solve_qs=0, solve_a_corr=0, precision=32,
      max_samples=320), syn, 300000)
check("mode C 1000x20x4x5", H.default_config(n_sta=20, n_events=1000, n_procs=4, n_chains=5, n_cool=1, n_iter=60000, n_burn=1000,
      n_interval=500, mode=H.MODE_BLOCKED_GIBBS, precision=32, max_samples=140), syn, 60000)
syn2 = H.Synthetic(7000, 20, 6)
check("mode C 7000x20x4x5 (persistent octet sweep)", H.default_config(n_sta=20, n_events=7000, n_procs=4, n_chains=5, n_cool=1,
      n_iter=6000, n_burn=100, n_interval=100, mode=H.MODE_BLOCKED_GIBBS, precision=32, max_samples=70), syn2, 6000)
