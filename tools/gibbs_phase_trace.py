"""Where one blocked-Gibbs iteration (float32 kernel) spends its time, from globaltimer stamps of CTA (0,0):
   python tools/gibbs_phase_trace.py E S iters [R K]              (one GPU)
   torchrun --nproc-per-node N tools/gibbs_phase_trace.py E S iters   (one ensemble, events sharded over N GPUs)
Needs the variant library built by `tools/build_variants.sh phase htm_gibbs_f32.cu "-DHTM_GIBBS_PHASE_TRACE"`."""
import ctypes
import os
import sys

sys.path.insert(0, ".")
os.environ.setdefault("HTM_B200_LIB", os.path.join(os.getcwd(), "variants", "libhtm_phase.so"))
import numpy as np

import hypotremormcmc_b200 as H

E, S, n_it = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
R, K = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (4, 5)
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
syn = H.Synthetic(E, S, 5)
kw = {}
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    syn = syn.shard(rank, world)
    kw = dict(device=local, shard_rank=rank, shard_count=world, gibbs_shard_events=1)
cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=10 ** 6, n_burn=0, n_interval=50,
                       mode=H.MODE_BLOCKED_GIBBS, precision=32, **kw)
with H.HypoTremorB200(cfg) as g:
    g.load(syn)
    g.init_chains()
    if world > 1:
        ids = [g.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        g.comm_init(ids[0])
        handles = [None] * world
        dist.all_gather_object(handles, g.comm_p2p_export())
        g.comm_p2p_import(handles)
    g.run(1, 20)
    g.synchronize()
    if world > 1:
        dist.barrier()
    g.run(21, 20 + n_it)
    g.synchronize()
    ms, _, _ = g.last_run_stats()
    n = min(n_it, 4096)
    buf = np.zeros((n, 8), dtype=np.uint64)
    rc = g.lib.htm_debug_phase_trace(buf.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), n)
    assert rc == 0
    cta = np.zeros((4096, 2), dtype=np.uint64)
    assert g.lib.htm_debug_cta_trace(cta.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), 4096) == 0
    info = np.zeros((4096, 4), dtype=np.uint32)
    assert g.lib.htm_debug_cta_info(info.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), 4096) == 0
if rank == 0:
    t = buf.astype(np.float64)[5:]                      # skip the first iterations
    sharded = world > 1
    names = ["stage proposals", "sweep (CTA 0,0)", "grid barrier 1"] + (["sum + peer exchange", "grid barrier 2"] if sharded else []) + ["decide"]
    cols = [(0, 1), (1, 2), (2, 3)] + ([(3, 4), (4, 5), (5, 6)] if sharded else [(3, 6)])
    total = np.median(t[1:, 0] - t[:-1, 0]) / 1e3
    print("E=%d S=%d J=%d on %d GPU(s): %.1f us per iteration (events: %.1f us); medians of CTA (0,0):"
          % (E, S, R * K, world, total, ms * 1e3 / n_it))
    for nm, (a, b) in zip(names, cols):
        print("  %-22s %7.2f us" % (nm, np.median(t[:, b] - t[:, a]) / 1e3))
    c = cta[cta[:, 1] > 0].astype(np.float64)
    if len(c) > 1:
        t0 = c[:, 0].min()
        dur, end = (c[:, 1] - c[:, 0]) / 1e3, (c[:, 1] - t0) / 1e3
        print("  sweep of one iteration over %d CTAs: duration min / median / max %.1f / %.1f / %.1f us; start spread %.1f us; "
              "last CTA ends %.1f us after the first starts" % (len(c), dur.min(), np.median(dur), dur.max(),
                                                               (c[:, 0].max() - t0) / 1e3, end.max()))
        inf = info[cta[:, 1] > 0].astype(np.int64)
        for nm, col in (("CTA row (blockIdx.y)", 1), ("event octets", 2)):
            for v in np.unique(inf[:, col]):
                k = inf[:, col] == v
                print("    %-22s %3d: %4d CTAs, %d warps, duration median %.1f, max %.1f us" % (nm, v, k.sum(), inf[k, 3][0], np.median(dur[k]), dur[k].max()))
        # what shares an SM: warps and octet visits per SM against the time its last CTA ends
        sm = inf[:, 0]
        ids = np.unique(sm)
        work = np.array([(inf[sm == i, 2] * inf[sm == i, 3]).sum() for i in ids])
        warps = np.array([inf[sm == i, 3].sum() for i in ids])
        last = np.array([end[sm == i].max() for i in ids])
        print("    per SM (%d SMs): CTAs %d..%d, warps %d..%d, warp-octets %d..%d; end of the SM's last CTA min / median / max %.1f / %.1f / %.1f us"
              % (len(ids), min((sm == i).sum() for i in ids), max((sm == i).sum() for i in ids), warps.min(), warps.max(), work.min(), work.max(),
                 last.min(), np.median(last), last.max()))
        for wv in np.unique(work):
            k = work == wv
            print("      SMs with %4d warp-octets (%d warps): %3d, end median %.1f max %.1f us" % (wv, warps[k][0], k.sum(), np.median(last[k]), last[k].max()))
        order = np.nonzero(cta[:, 1] > 0)[0]
        print("      SM of CTAs 0..23: %s" % " ".join(str(v) for v in sm[:24]))
        for i in ids[:4]:
            print("      SM %d holds CTAs %s (ends %s us)" % (i, list(order[sm == i]), ["%.0f" % v for v in end[sm == i]]))
        half = ids < np.median(ids)
        print("      SM id below / above the median id: end median %.1f / %.1f us" % (np.median(last[half]), np.median(last[~half])))
        print("      correlation(end, warp-octets) = %.2f" % np.corrcoef(last, work)[0, 1])
if world > 1:
    dist.destroy_process_group()
