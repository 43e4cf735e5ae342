"""Generates tests/golden/schedule_golden.npz: short trajectories of the CPU oracle's three schedules (mode A =
the reference's own, fed by mod_random; mode B factorised and mode C blocked Gibbs on Philox draws) for fixed
seeds (flags at every step, log-likelihoods at every 10th iteration).  These are REGRESSION pins: they freeze the schedules (draw order, proposal mix, judge, swap, recording)
as they were when the oracle was checked against the independent Python restatement and the statistical tests,
so that a later refactor cannot silently change every chain.  The oracle itself remains "parity unpinned" by
the reference (no Fortran compiler here, the reference ships no vectors).

    python tests/golden/make_schedule_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import hypotremormcmc_b200 as H  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402

CASES = {
    "A": dict(E=3, S=7, seed=31, cfg=dict(n_procs=2, n_chains=3, n_cool=1, n_iter=400, n_burn=50, n_interval=10,
                                          mode=H.MODE_REPLAY, precision=64)),
    "B": dict(E=4, S=9, seed=32, cfg=dict(n_procs=2, n_chains=4, n_cool=1, n_iter=300, n_burn=50, n_interval=10,
                                          mode=H.MODE_FACTORISED, precision=64, solve_vs=0, solve_t_corr=0, solve_qs=0,
                                          solve_a_corr=0)),
    "C": dict(E=4, S=9, seed=33, cfg=dict(n_procs=2, n_chains=3, n_cool=1, n_iter=300, n_burn=50, n_interval=10,
                                          mode=H.MODE_BLOCKED_GIBBS, precision=64)),
}


def run_case(c):
    syn = H.Synthetic(c["E"], c["S"], c["seed"])
    cfg = H.default_config(n_sta=c["S"], n_events=c["E"], **c["cfg"])
    o = Oracle(cfg, syn)
    o.init_chains()
    tr, sw = o.run(1, cfg.n_iter)
    p, a = o.get_counts()
    smp = o.fetch_samples(0)
    return dict(proposal_type=tr["proposal_type"].astype(np.int8), index=tr["index"].astype(np.int32),
                prior_ok=tr["prior_ok"].astype(np.int8), accepted=tr["accepted"].astype(np.int8),
                log_likelihood=tr["log_likelihood"][9::10], swaps=np.asarray(sw), n_propose=p, n_accept=a,
                sample_iter=smp["iter"], sample_hypo=smp["hypo"])


if __name__ == "__main__":
    out = {}
    for name, c in CASES.items():
        for k, v in run_case(c).items():
            out["%s_%s" % (name, k)] = v
    np.savez_compressed(os.path.join(HERE, "schedule_golden.npz"), **out)
    print("wrote schedule_golden.npz:", {k: v.shape for k, v in out.items() if k.endswith("accepted")})
