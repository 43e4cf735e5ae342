"""Multi-GPU check of the event-sharded blocked-Gibbs mode (run under torchrun, one rank per GPU): the events of
every joint chain are split over the ranks, each iteration all-reduces the per-chain sums (NCCL inside the
library), every rank takes the same decisions.  Compared step by step with the oracle's UNSHARDED statement."""
import os
import sys

sys.path.insert(0, ".")
import numpy as np
import torch
import torch.distributed as dist

import hypotremormcmc_b200 as H
from oracle.pyoracle import Oracle

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
E, S, R, K, n_it = 203, 20, 2, 4, 60
syn = H.Synthetic(E, S, 9)
base = dict(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=n_it, n_burn=10, n_interval=5,
            mode=H.MODE_BLOCKED_GIBBS, precision=64, max_samples=16)
ids = [H.HypoTremorB200.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
cfg = H.default_config(device=local, shard_rank=rank, shard_count=world, gibbs_shard_events=1, **base)
sh = syn.shard(rank, world)
with H.HypoTremorB200(cfg) as g:
    assert g.n_events == sh.n_events and g.n_procs == R
    g.load(sh)
    g.init_chains()
    g.comm_init(ids[0])
    if os.environ.get("HTM_GIBBS_EXCHANGE", "p2p") != "nccl":  # per-iteration exchange through NVLink peer memory
        mine = g.comm_p2p_export()
        handles = [None] * world
        dist.all_gather_object(handles, mine)
        g.comm_p2p_import(handles)
    tr, sw = g.run_traced(1, n_it)
    st = g.get_chain_state(1, 2)
    _, p, a = g.gather(histograms=False)
    smp = g.fetch_samples(0)
o = Oracle(H.default_config(**base), syn)
o.init_chains()
tr_o, sw_o = o.run(1, n_it)
lo = sh.event_offset
ok = True
for f in ("proposal_type", "index", "prior_ok", "accepted"):
    hyp = np.array_equal(tr[f][:, :-1], tr_o[f][:, lo:lo + sh.n_events])
    if f == "index":   # event ids inside the trace are shard-local
        hyp = np.array_equal(tr[f][:, :-1] + 3 * lo, tr_o[f][:, lo:lo + sh.n_events])
    ok &= hyp and np.array_equal(tr[f][:, -1], tr_o[f][:, -1])
ok &= bool(np.array_equal(sw, sw_o))
relL = np.max(np.abs(tr["log_likelihood"][:, -1] - tr_o["log_likelihood"][:, -1]) / np.abs(tr_o["log_likelihood"][:, -1]))
so = o.get_chain_state(1, 2)
ok &= relL < 1e-9 and abs(st["vs"] - so["vs"]) < 1e-12 and st["temp"] == so["temp"]
ok &= bool(np.allclose(st["hypo"], so["hypo"][3 * lo:3 * (lo + sh.n_events)], rtol=1e-10, atol=1e-10))
po, ao = o.get_counts()
ok &= bool(np.array_equal(p, po) and np.array_equal(a, ao))
os_ = o.fetch_samples(0)
ok &= bool(np.array_equal(smp["iter"], os_["iter"]) and np.allclose(smp["t_corr"], os_["t_corr"], atol=1e-12))
print("comm_check_gibbs rank %d/%d: matches the unsharded oracle: %s (max rel dL %.2e)" % (rank, world, ok, relL))
flag = torch.tensor([1 if ok else 0])
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(flag) == 1 else 1)
