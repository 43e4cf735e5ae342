"""Event-sharded blocked Gibbs in float32 under torchrun (one rank per GPU): the sharded run (persistent kernel with the
peer-memory exchange between two grid barriers) must reproduce, flag for flag and bit for bit, the UNSHARDED float32
run of the same library on one GPU (which the single-GPU suite ties to the oracle).

HTM_TEST_DELAY_S=<s>: the last rank starts its run <s> seconds late -- the exchange must simply wait (wall-clock
budget HTM_XCH_TIMEOUT_S, default 120 s).  With HTM_TEST_EXPECT_TIMEOUT=1 the delay exceeds the budget: every rank
must then get an ERROR from the calls that return results, never numbers decided on partial sums."""
import os
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch
import torch.distributed as dist

import hypotremormcmc_b200 as H

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
E, S, R, K, n_it = 2011, 20, 2, 4, 80
syn = H.Synthetic(E, S, 9)
base = dict(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=n_it, n_burn=10, n_interval=5,
            mode=H.MODE_BLOCKED_GIBBS, precision=32, max_samples=32)
cfg = H.default_config(device=local, shard_rank=rank, shard_count=world, gibbs_shard_events=1, **base)
sh = syn.shard(rank, world)
ids = [H.HypoTremorB200.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
with H.HypoTremorB200(cfg) as g:
    g.load(sh)
    g.init_chains()
    g.comm_init(ids[0])
    if os.environ.get("HTM_GIBBS_EXCHANGE", "p2p") != "nccl":
        mine = g.comm_p2p_export()
        handles = [None] * world
        dist.all_gather_object(handles, mine)
        g.comm_p2p_import(handles)
    delay = float(os.environ.get("HTM_TEST_DELAY_S", "0"))
    if delay and rank == world - 1:
        time.sleep(delay)
    if os.environ.get("HTM_TEST_EXPECT_TIMEOUT"):
        failed = []
        for call in (lambda: g.run_traced(1, 50), g.get_counts, lambda: g.get_chain_state(0, 0), lambda: g.fetch_samples(0),
                     g.synchronize):
            try:
                call()
                failed.append(False)
            except H.HtmError as ex:
                failed.append("exchange timed out" in str(ex))
        print("comm_check_gibbs_f32 rank %d/%d: every result call reports the exchange time-out: %s" % (rank, world, all(failed)))
        flag = torch.tensor([1 if all(failed) else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        dist.destroy_process_group()
        sys.exit(0 if int(flag) == 1 else 1)
    tr, sw = g.run_traced(1, 50)
    g.run(51, n_it)
    _, nl, _ = g.last_run_stats()
    st = g.get_chain_state(1, 2)
    _, p, a_ = g.gather(histograms=False)
    smp = g.fetch_samples(0)
with H.HypoTremorB200(H.default_config(device=local, **base)) as u:  # the same chains, unsharded, on this GPU
    u.load(syn)
    u.init_chains()
    tr_u, sw_u = u.run_traced(1, 50)
    u.run(51, n_it)
    su = u.get_chain_state(1, 2)
    pu, au = u.get_counts()
    smp_u = u.fetch_samples(0)
lo = sh.event_offset
# What must be IDENTICAL: every decision (flags, swaps), hence every state, counter and record.  The per-chain sums
# over all events are float64 accumulations of float32 terms of very different magnitude (L_e and the per-event
# change dL_e of the pending proposal), so their last bits depend on how the events are grouped into partial sums:
# sharded and unsharded totals agree to rounding (1e-13 relative), and on every shard they are the same bits.
checks = {}
for f in ("proposal_type", "prior_ok", "accepted"):
    checks[f] = bool(np.array_equal(tr[f][:, :-1], tr_u[f][:, lo:lo + sh.n_events]) and np.array_equal(tr[f][:, -1], tr_u[f][:, -1]))
checks["swaps"] = bool(np.array_equal(sw, sw_u))
a, b = tr["log_likelihood"][:, -1], tr_u["log_likelihood"][:, -1]
checks["sums"] = bool(np.all(np.abs(a - b) <= 1e-13 * np.abs(b)))
checks["shared"] = bool(st["vs"] == su["vs"] and st["qs"] == su["qs"] and st["temp"] == su["temp"] and
                        abs(st["log_likelihood"] - su["log_likelihood"]) <= 1e-13 * abs(su["log_likelihood"]))
checks["state"] = bool(np.array_equal(st["hypo"], su["hypo"][3 * lo:3 * (lo + sh.n_events)]) and np.array_equal(st["t_corr"], su["t_corr"]))
checks["counts"] = bool(np.array_equal(p, pu) and np.array_equal(a_, au))
checks["records"] = bool(np.array_equal(smp["iter"], smp_u["iter"]) and np.array_equal(smp["vs"], smp_u["vs"]))
# every shard holds the same bits: compare the summed log-likelihood trace across ranks
mine = torch.from_numpy(np.ascontiguousarray(a))
lo_t, hi_t = mine.clone(), mine.clone()
dist.all_reduce(lo_t, op=dist.ReduceOp.MIN)
dist.all_reduce(hi_t, op=dist.ReduceOp.MAX)
checks["replicated_bits"] = bool(torch.equal(lo_t, hi_t))
ok = all(checks.values())
if not ok:
    print("comm_check_gibbs_f32 rank %d: %r" % (rank, checks))
print("comm_check_gibbs_f32 rank %d/%d (%s exchange, %d launches in the last run): equals the unsharded run: %s"
      % (rank, world, os.environ.get("HTM_GIBBS_EXCHANGE", "p2p"), nl, ok))
flag = torch.tensor([1 if ok else 0])
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(flag) == 1 else 1)
