"""N > 1 host logic on CPU: two gloo ranks each run THEIR shard of events (the oracle's statement of
the factorised schedule stands in for the GPU) and gather histograms / reduce counters with the same
code the GPU run uses; the result must equal the unsharded run."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import hypotremormcmc_b200 as H
from hypotremormcmc_b200.gather import gather_event_blocks, reduce_counts

E, S, R, K, BINS, N_IT = 7, 6, 2, 3, 8, 120


def cfg_of(**kw):
    return H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=N_IT, n_burn=20,
                            n_interval=4, mode=H.MODE_FACTORISED, precision=64, solve_vs=0, solve_t_corr=0,
                            solve_qs=0, solve_a_corr=0, hist_bins=BINS, **kw)


def histogram_of(samples, syn, cfg):
    """the device histogram rule (csrc/htm_factorised.cu: hist_add) on fetched samples"""
    h = np.zeros((syn.n_events, 3, BINS), dtype=np.int32)
    for rec in samples:
        xyz = rec.reshape(-1, 3)
        for e in range(syn.n_events):
            bx = int(np.floor((xyz[e, 0] - syn.x_mu[e] + cfg.hist_xy_halfwidth) * BINS / (2 * cfg.hist_xy_halfwidth)))
            by = int(np.floor((xyz[e, 1] - syn.y_mu[e] + cfg.hist_xy_halfwidth) * BINS / (2 * cfg.hist_xy_halfwidth)))
            bz = int(np.floor((xyz[e, 2] - cfg.prior_z) * BINS / cfg.hist_z_max))
            for c, b in enumerate((bx, by, bz)):
                h[e, c, min(max(b, 0), BINS - 1)] += 1
    return h


def run_shard(syn, cfg, offset):
    from oracle.pyoracle import Oracle
    o = Oracle(cfg, syn, event_offset=offset)
    o.init_chains()
    o.run(1, N_IT, trace=False)
    samples = np.concatenate([o.fetch_samples(r)["hypo"] for r in range(R)])
    p, a = o.get_counts()
    return histogram_of(samples, syn, cfg), np.concatenate([p, a])


def worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    syn = H.Synthetic(E, S, 77)
    sh = syn.shard(rank, world)
    hist, counts = run_shard(sh, cfg_of(shard_rank=rank, shard_count=world), sh.event_offset)
    full = gather_event_blocks(torch.from_numpy(hist), E)
    tot = reduce_counts(torch.from_numpy(counts))
    if rank == 0:
        out["hist"] = full.numpy().copy()
        out["counts"] = tot.numpy().copy()
    dist.barrier()
    dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_gather_equals_unsharded(world):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(worker, args=(world, free_port(), out), nprocs=world, join=True)
    syn = H.Synthetic(E, S, 77)
    hist, counts = run_shard(syn, cfg_of(), 0)
    assert np.array_equal(out["hist"], hist)
    assert np.array_equal(out["counts"], counts)
    assert hist.sum() == 3 * E * R * ((N_IT - 20) // 4)


def test_shard_bounds_cover_everything():
    for n in (1, 7, 1000, 100000):
        for w in (1, 2, 3, 8):
            if w > n:
                continue
            edges = [H.shard_bounds(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
