"""Multi-GPU check of the ABI-level gather (run under torchrun, one rank per GPU):
every rank runs ITS shard of events, then htm_gather (NCCL inside the library) must reproduce the
histograms and counters of the unsharded run that rank 0 also does, and htm_gather_samples must deliver, on
every rank, the records of each virtual rank with the hypocentres of ALL events."""
import os
import sys

sys.path.insert(0, ".")
import numpy as np
import torch
import torch.distributed as dist

import hypotremormcmc_b200 as H

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")     # only to hand the NCCL id around (the host program's job)
E, S, R, K = 1003, 20, 4, 16
syn = H.Synthetic(E, S, 9)
base = dict(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=300, n_burn=0, n_interval=10,
            mode=H.MODE_FACTORISED, precision=32, hist_bins=32, max_samples=40, solve_vs=0, solve_t_corr=0, solve_qs=0, solve_a_corr=0)
ids = [H.HypoTremorB200.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
cfg = H.default_config(device=local, shard_rank=rank, shard_count=world, **base)
with H.HypoTremorB200(cfg) as g:
    g.load(syn.shard(rank, world))
    g.init_chains()
    g.run(1, 300)
    g.comm_init(ids[0])
    hist, p, a = g.gather()
    smp = [g.gather_samples(r) for r in range(R)]
ok_s = True
with H.HypoTremorB200(H.default_config(device=local, **base)) as g:  # the unsharded run, on every rank
    g.load(syn)
    g.init_chains()
    g.run(1, 300)
    for r in range(R):
        s0 = g.fetch_samples(r)
        ok_s &= bool(np.array_equal(s0["iter"], smp[r]["iter"]) and np.array_equal(s0["hypo"], smp[r]["hypo"]))
        ok_s &= len(s0["iter"]) == 30 and smp[r]["hypo"].shape == (30, 3 * E)
print("comm_check rank %d/%d: gathered samples == unsharded samples: %s" % (rank, world, ok_s))
assert ok_s
if rank == 0:
    with H.HypoTremorB200(H.default_config(device=local, **base)) as g:
        g.load(syn)
        g.init_chains()
        g.run(1, 300)
        h0 = g.get_histograms()
        p0, a0 = g.get_counts()
    ok = np.array_equal(hist, h0) and np.array_equal(p, p0) and np.array_equal(a, a0)
    print("comm_check world=%d: gathered == unsharded: %s (hist sum %d, cold proposals %d)" % (world, ok, hist.sum(), p.sum()))
    assert ok
dist.barrier()
dist.destroy_process_group()
