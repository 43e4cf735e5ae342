"""The C++ oracle against a second, independently structured pure-Python restatement of the reference's
whole main loop (tests/pyref_mode_a.py): every proposal, judgement, swap, counter and recorded sample."""
import numpy as np
import pytest

import hypotremormcmc_b200 as H
from oracle.pyoracle import Oracle

import pyref_mode_a as pyref


@pytest.mark.parametrize("E,S,R,K,n_cool,solve,use", [(2, 5, 2, 3, 1, (1, 1, 1, 1), (1, 1)), (1, 4, 3, 2, 1, (0, 0, 0, 0), (1, 1)),
                                                       (3, 3, 1, 4, 2, (1, 0, 0, 1), (1, 0)), (2, 6, 2, 2, 1, (0, 1, 1, 0), (0, 1))])
def test_oracle_equals_python_restatement(E, S, R, K, n_cool, solve, use):
    syn = H.Synthetic(E, S, 300 + E + S)
    syn.t_stdv[0, 1] = 0.0                      # exercises the degenerate-sigma rule inside the loop too
    n_it = 260
    cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=n_cool, n_iter=n_it, n_burn=60, n_interval=7,
                           mode=H.MODE_REPLAY, precision=64, solve_vs=solve[0], solve_t_corr=solve[1], solve_qs=solve[2],
                           solve_a_corr=solve[3], use_time=use[0], use_amp=use[1])
    o = Oracle(cfg, syn)
    o.init_chains()
    tr, sw = o.run(1, n_it)
    ptrace, pswaps, psamples, pcounts = pyref.run(syn, cfg, syn.x_mu, syn.y_mu, n_it)
    pt = np.array([t[:4] for t in ptrace]).reshape(n_it, R, K, 4)
    pl = np.array([t[4] for t in ptrace]).reshape(n_it, R, K)
    assert np.array_equal(tr["proposal_type"], pt[..., 0]) and np.array_equal(tr["index"], pt[..., 1])
    assert np.array_equal(tr["prior_ok"], pt[..., 2]) and np.array_equal(tr["accepted"], pt[..., 3])
    assert np.allclose(tr["log_likelihood"], pl, rtol=1e-12, atol=0)
    ps = np.array(pswaps)
    for k, f in enumerate(("rank1", "chain1", "rank2", "chain2", "accepted")):
        assert np.array_equal(sw[f], ps[:, k]), f
    co = o.get_counts()
    assert list(co[0]) == pcounts[0] and list(co[1]) == pcounts[1]
    for r in range(R):
        s = o.fetch_samples(r)
        assert list(s["iter"]) == [x[0] for x in psamples[r]]
        if len(psamples[r]):
            assert np.allclose(s["vs"], [x[1] for x in psamples[r]], rtol=1e-13)
            assert np.allclose(s["hypo"], np.array([x[2] for x in psamples[r]]), rtol=1e-12, atol=1e-12)
            assert np.allclose(s["t_corr"], np.array([x[3] for x in psamples[r]]), rtol=1e-12, atol=1e-13)
            assert np.allclose(s["a_corr"], np.array([x[5] for x in psamples[r]]), rtol=1e-12, atol=1e-13)
