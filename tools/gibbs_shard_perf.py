"""Throughput of the event-sharded blocked-Gibbs mode under torchrun (one rank per GPU)."""
import os
import sys
import time

sys.path.insert(0, ".")
import torch
import torch.distributed as dist

import hypotremormcmc_b200 as H

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
E, S, R, K, n_it = int(sys.argv[1]), int(sys.argv[2]), 4, 5, int(sys.argv[3])
syn = H.Synthetic(E, S, 5).shard(rank, world)
ids = [H.HypoTremorB200.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=n_it, n_burn=0, n_interval=50,
                       mode=H.MODE_BLOCKED_GIBBS, precision=32, device=local, shard_rank=rank, shard_count=world,
                       gibbs_shard_events=1)
with H.HypoTremorB200(cfg) as g:
    g.load(syn)
    g.init_chains()
    g.comm_init(ids[0])
    if os.environ.get("HTM_GIBBS_EXCHANGE", "p2p") != "nccl":  # per-iteration exchange through NVLink peer memory
        mine = g.comm_p2p_export()
        handles = [None] * world
        dist.all_gather_object(handles, mine)
        g.comm_p2p_import(handles)
    g.run(1, 10)
    g.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    g.run(11, 10 + n_it)
    g.synchronize()
    dt = time.perf_counter() - t0
    ms, nl, npr = g.last_run_stats()
t = torch.tensor([dt])
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("event-sharded blocked Gibbs (%s exchange), %d GPUs, E=%d S=%d J=%d: %.1f us/iteration, %.3g proposals/s (all shards)"
          % (os.environ.get("HTM_GIBBS_EXCHANGE", "p2p"), world, E, S, R * K, float(t) * 1e6 / n_it, n_it * (E + 1) * R * K / float(t)))
dist.destroy_process_group()
