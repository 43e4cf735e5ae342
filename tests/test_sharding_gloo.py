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


# ---- blocked Gibbs with the events of every joint chain sharded (the one exchange step of the path) --------------
GE, GS, GR, GK, G_IT = 9, 6, 2, 3, 80


def gibbs_cfg(**kw):
    return H.default_config(n_sta=GS, n_events=GE, n_procs=GR, n_chains=GK, n_cool=1, n_iter=G_IT, n_burn=10,
                            n_interval=5, mode=H.MODE_BLOCKED_GIBBS, precision=64, **kw)


def gibbs_worker(rank, world, port, out):
    """One rank = one event shard of ALL joint chains.  Per chain and iteration the two sums over events (current,
    proposed) are all-gathered and added in shard order -- what peer_allreduce (csrc/htm_gibbs_decide.cuh) does on
    the device -- so every rank takes the same decisions."""
    from oracle.pyoracle import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    syn = H.Synthetic(GE, GS, 78)
    sh = syn.shard(rank, world)
    o = Oracle(gibbs_cfg(shard_rank=rank, shard_count=world, gibbs_shard_events=1), sh, event_offset=sh.event_offset)
    o.init_chains()

    def exchange(cur, prop):
        mine = torch.tensor([cur, prop], dtype=torch.float64)
        parts = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, mine)
        a = b = 0.0
        for p in parts:  # shard order: the same operands in the same order on every rank
            a += float(p[0])
            b += float(p[1])
        return a, b

    o.set_sum_hook(exchange)
    tr, sw = o.run(1, G_IT)
    out[rank] = dict(lo=sh.event_offset, n=sh.n_events, tr={k: np.array(tr[k]) for k in tr.dtype.names}, sw=np.array(sw),
                     state=o.get_chain_state(1, 1), samples=o.fetch_samples(0), lik=o.fetch_likelihood(0))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_event_sharded_blocked_gibbs_equals_unsharded(world):
    from oracle.pyoracle import Oracle
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(gibbs_worker, args=(world, free_port(), out), nprocs=world, join=True)
    syn = H.Synthetic(GE, GS, 78)
    u = Oracle(gibbs_cfg(), syn)
    u.init_chains()
    tr_u, sw_u = u.run(1, G_IT)
    st_u, smp_u, lik_u = u.get_chain_state(1, 1), u.fetch_samples(0), u.fetch_likelihood(0)
    covered = 0
    for rank in range(world):
        r = out[rank]
        lo, n = r["lo"], r["n"]
        covered += n
        for f in ("proposal_type", "prior_ok", "accepted"):
            assert np.array_equal(r["tr"][f][:, :n], tr_u[f][:, lo:lo + n]), (rank, f)      # this shard's events
            assert np.array_equal(r["tr"][f][:, n], tr_u[f][:, GE]), (rank, f)               # shared-parameter row
        assert np.array_equal(r["tr"]["index"][:, :n] + 3 * lo, tr_u["index"][:, lo:lo + n])  # indices are shard-local
        assert np.array_equal(r["sw"], sw_u)                                                # every rank: the same swaps
        assert np.allclose(r["tr"]["log_likelihood"][:, n], tr_u["log_likelihood"][:, GE], rtol=1e-12, atol=0)
        assert abs(r["state"]["vs"] - st_u["vs"]) < 1e-13 and r["state"]["temp"] == st_u["temp"]
        assert np.allclose(r["state"]["hypo"], st_u["hypo"][3 * lo:3 * (lo + n)], rtol=1e-12, atol=1e-12)
        assert np.array_equal(r["samples"]["iter"], smp_u["iter"])
        assert np.allclose(r["samples"]["hypo"], smp_u["hypo"][:, 3 * lo:3 * (lo + n)], rtol=1e-12, atol=1e-12)
        assert np.allclose(r["samples"]["t_corr"], smp_u["t_corr"], atol=1e-13)
        assert np.array_equal(r["lik"][0], lik_u[0]) and np.allclose(r["lik"][1], lik_u[1], rtol=1e-12)
    assert covered == GE
