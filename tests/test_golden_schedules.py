"""Regression pins of the three schedules (tests/golden/schedule_golden.npz, made by
tests/golden/make_schedule_golden.py): the CPU oracle must still produce them, and -- on a GPU -- so must the
float64 CUDA kernels of modes B and C through the C ABI."""
import os
import sys

import numpy as np
import pytest

import hypotremormcmc_b200 as H
from oracle.pyoracle import Oracle

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_schedule_golden import CASES, run_case  # noqa: E402

G = np.load(os.path.join(HERE, "golden", "schedule_golden.npz"))
FLAGS = ("proposal_type", "index", "prior_ok", "accepted")


@pytest.mark.parametrize("name", ["A", "B", "C"])
def test_oracle_reproduces_the_golden_trajectories(name):
    got = run_case(CASES[name])
    for k in FLAGS + ("swaps", "n_propose", "n_accept", "sample_iter"):
        assert np.array_equal(got[k], G["%s_%s" % (name, k)]), k
    assert np.allclose(got["log_likelihood"], G[name + "_log_likelihood"], rtol=1e-12, atol=0)
    assert np.allclose(got["sample_hypo"], G[name + "_sample_hypo"], rtol=1e-12, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["B", "C"])
def test_cuda_float64_reproduces_the_golden_trajectories(name):
    c = CASES[name]
    syn = H.Synthetic(c["E"], c["S"], c["seed"])
    cfg = H.default_config(n_sta=c["S"], n_events=c["E"], max_samples=64, **c["cfg"])
    with H.HypoTremorB200(cfg) as g:
        g.load(syn)
        g.init_chains()
        tr, sw = g.run_traced(1, cfg.n_iter)
        p, a = g.get_counts()
        smp = g.fetch_samples(0)
    for k in FLAGS:
        assert np.array_equal(tr[k].reshape(G["%s_%s" % (name, k)].shape), G["%s_%s" % (name, k)]), k
    assert np.array_equal(np.asarray(sw).reshape(G[name + "_swaps"].shape), G[name + "_swaps"])
    assert np.array_equal(p, G[name + "_n_propose"]) and np.array_equal(a, G[name + "_n_accept"])
    L = tr["log_likelihood"].reshape((cfg.n_iter,) + G[name + "_log_likelihood"].shape[1:])[9::10]
    assert np.allclose(L, G[name + "_log_likelihood"], rtol=1e-9, atol=0)
    assert np.array_equal(smp["iter"], G[name + "_sample_iter"])
    assert np.allclose(smp["hypo"], G[name + "_sample_hypo"], rtol=1e-10, atol=1e-10)
