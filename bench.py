#!/usr/bin/env python
"""bench.py -- MCMC proposals/s of the hypo_tremor_mcmc hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[1]): 1,000 synthetic events x 20 stations x 16 temperatures x
4 chains per GPU = 64,000 tempered chains, factorised mode (solve_* = F), float32, Philox.
One "step" = `--iters` Metropolis iterations of every chain (one kernel launch).  With N > 1
every rank owns its own 1,000 events (events shard; no data-path collective) -> weak scaling.

Numbers on the JSON line:
  value     proposals/s, inputs resident in HBM, CUDA-event time of the K timed steps, max over ranks
  e2e       the same metric through the C ABI with HOST buffers: each step is a whole batch job
            (htm_set_observations H2D -> htm_init_chains -> htm_run -> fetch samples, likelihood,
            histograms, counts D2H), host wall clock
  roofline  achieved algorithmic FLOP/s ((30*S+64) per proposal) of the
            dominant kernel / own-measured FFMA peak (no driver-measured FP32 peak exists);
            roofline.traffic = DRAM bytes per launch from the committed ncu capture (default workload only)
            and roofline.hbm = that traffic per launch time against MEASURED_PEAKS.json's copy bandwidth
            (the evidence that the kernel is not HBM-bound)
  cpu_baseline  the C++ restatement of the reference algorithm (oracle/, mode A: joint chain, one
            scalar per iteration, one swap per iteration), one thread per virtual rank, on a
            bounded sample of the same workload.  The real Fortran/MPI binary cannot be built in
            this image (no gfortran/mpif90), so kind = "port".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MCMC proposals/sec (events x chains)"
UNIT = "proposals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--events", type=int, default=1000, help="events per GPU")
    ap.add_argument("--stations", type=int, default=20)
    ap.add_argument("--ranks", type=int, default=4, help="virtual ranks = independent tempering groups per event")
    ap.add_argument("--chains", type=int, default=16, help="chains (temperatures) per rank")
    ap.add_argument("--iters", type=int, default=20000, help="iterations per step")
    ap.add_argument("--interval", type=int, default=1000)
    ap.add_argument("--precision", type=int, default=32, choices=[32, 64])
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 warp-per-chain, 2 lane-per-chain")
    ap.add_argument("--slots", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target duration of the CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def workload_name(a, n_gpus):
    shape = (a.events, a.stations, a.chains, a.ranks)
    tag = {(1000, 20, 16, 4): "BASELINE configs[1]", (10000, 50, 16, 4): "events x stations of BASELINE configs[2]",
           (12500, 50, 16, 4): "BASELINE configs[3] when run on 8 GPUs"}.get(shape, "not a BASELINE config")
    return "%d synthetic events x %d stations x %d temperatures x %d chains per GPU (%s)" % (shape + (tag,))


def make_cfg(H, a, n_events_total, shard_rank, shard_count, device, max_samples, hist_bins):
    return H.default_config(
        n_sta=a.stations, n_events=n_events_total, n_procs=a.ranks, n_chains=a.chains, n_cool=1,
        n_iter=a.iters, n_burn=0, n_interval=a.interval, mode=H.MODE_FACTORISED,
        solve_vs=0, solve_t_corr=0, solve_qs=0, solve_a_corr=0, precision=a.precision,
        ladder=H.LADDER_RANDOM, kernel=a.kernel, lane_slots=a.slots, device=device,
        shard_rank=shard_rank, shard_count=shard_count, max_samples=max_samples, hist_bins=hist_bins,
        seed=20231002)


# ---------------------------------------------------------------------------------------------
# clocks sampled DURING the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled through NVML every 20 ms by a thread."""

    def __init__(self, device):
        self.device, self.sm, self.reasons, self.stop_flag, self.max_mhz, self.err = device, [], set(), False, None, None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            # LOCAL_RANK indexes CUDA_VISIBLE_DEVICES; map through the UUID-free common case
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.device
            if vis:
                try:
                    idx = int(vis.split(",")[self.device])
                except (ValueError, IndexError):
                    idx = self.device
            self.h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception as ex:  # no NVML: report that, do not fail the bench
            self.err = "nvml unavailable: %r" % (ex,)
            return
        self.th = threading.Thread(target=self._loop, daemon=True)
        self.th.start()

    def _loop(self):
        nv = self.nv
        bits = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                "hw_power_brake_slowdown": 0x80, "sw_power_cap": 0x4}
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, b in bits.items():
                    if r & b:
                        self.reasons.add(name)
            except Exception as ex:
                self.err = repr(ex)
                break
            time.sleep(0.02)

    def stop(self):
        self.stop_flag = True
        if self.err and not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "samples": 0, "reasons": [self.err]}
        self.th.join(timeout=1.0)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "samples": len(self.sm), "reasons": sorted(self.reasons)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle (C++ restatement of the reference algorithm), mode A
# ---------------------------------------------------------------------------------------------
def cpu_reference_rate(a, seconds, n_events=None):
    """proposals/s of the reference algorithm (joint chains, src/hypo_tremor_mcmc.f90:236-284) on the
    host cores: same stations/chains, solve_* = F like the GPU arm, one thread per virtual rank."""
    import hypotremormcmc_b200 as H
    from oracle.pyoracle import Oracle
    cores = os.cpu_count() or 1
    total = a.ranks * a.chains
    # as many virtual ranks (threads) as cores allow, keeping at least one hot chain per rank
    ranks = max(d for d in range(1, total + 1) if total % d == 0 and d <= cores and total // d >= 2)
    chains = total // ranks
    E = n_events or a.events
    syn = H.Synthetic(E, a.stations, 20231002)
    cfg = H.default_config(n_sta=a.stations, n_events=E, n_procs=ranks, n_chains=chains, n_cool=1,
                           n_iter=10 ** 9, n_burn=10 ** 9, n_interval=1000, mode=H.MODE_REPLAY, precision=64,
                           solve_vs=0, solve_t_corr=0, solve_qs=0, solve_a_corr=0)
    o = Oracle(cfg, syn)
    o.init_chains()
    # iteration 1 is a full O(E*S) likelihood per chain; time the steady state after it
    o.run_threaded(1, 2)
    it, n, wall = 3, 2000, 0.0
    done = 0
    while wall < seconds:
        t = o.run_threaded(it, it + n - 1)
        wall += t
        done += n
        it += n
        if t < 1.0:
            n *= 2
    rate = done * total / wall
    sample = ("oracle mode A (reference semantics), %d events x %d stations, %d virtual ranks x %d chains, "
              "%d iterations in %.1f s, %d threads" % (E, a.stations, ranks, chains, done, wall, ranks))
    o.close()
    return rate, ranks, sample


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle
    pyoracle.build()
    # each step a bounded sample; the whole --steps K --warmup W run stays within ~2 minutes
    per_step = min(a.cpu_seconds, max(1.0, min(20.0, 120.0 / max(1, a.steps + a.warmup))))
    rates = []
    sample, cores = "", 1
    for i in range(a.warmup + a.steps):
        # same config as the B200 arm at this N: weak scaling, a.events per GPU
        r, cores, sample = cpu_reference_rate(a, per_step, n_events=a.events * max(1, a.gpus))
        if i >= a.warmup:
            rates.append(r)
    v = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a, a.gpus), "mode": "A (reference joint chain)", "solve": "F",
                   "note": "C++ restatement of the reference algorithm; the Fortran/MPI binary cannot be built here"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
def run_b200(a):
    import torch
    import torch.distributed as dist
    import hypotremormcmc_b200 as H
    from hypotremormcmc_b200.api import measure_fp32_peak

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    E_total = a.events * world
    syn_all = H.Synthetic(E_total, a.stations, 20231002)
    syn = syn_all.shard(rank, world)
    n_rec = (a.iters + a.interval - 1) // a.interval + 1
    hist_bins = 64
    cfg = make_cfg(H, a, E_total, rank, world, local, n_rec, hist_bins)

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm: value ----------------------------------------------------------
    g = H.HypoTremorB200(cfg)
    g.load(syn)
    g.init_chains()
    g.synchronize()
    peak_tf, mufu = measure_fp32_peak(local)
    it0 = 1
    for _ in range(a.warmup):
        g.run(it0, it0 + a.iters - 1)
        g.synchronize()
        g.discard_samples()
        it0 += a.iters
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    t0 = time.perf_counter()
    ms_steps, launches = [], 0
    for _ in range(a.steps):
        flush.zero_()  # L2 flush between timed steps (outside the CUDA-event pair)
        torch.cuda.synchronize()
        g.run(it0, it0 + a.iters - 1)
        g.synchronize()
        ms, nl, npr = g.last_run_stats()
        ms_steps.append(ms)
        launches += nl
        g.discard_samples()
        it0 += a.iters
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    clk = clocks.stop()
    barrier()
    dev_ms = float(np.sum(ms_steps))
    proposals_per_rank = a.steps * a.iters * syn.n_events * a.ranks * a.chains
    p, acc = g.get_counts()

    # ---- end-to-end arm through the C ABI with host buffers ----------------------------------
    h2d = 4 * syn.t_obs.nbytes + syn.x_mu.nbytes * 2 + 3 * syn.sta_x.nbytes
    d2h = 0
    e2e_n = max(3, min(a.steps, 10))
    for i in range(2 + e2e_n):
        if i == 2:
            barrier()
            te = time.perf_counter()
        g.set_stations(syn.sta_x, syn.sta_y, syn.sta_z)
        g.set_observations(syn.t_obs, syn.t_stdv, syn.a_obs, syn.a_stdv)
        g.set_xy_prior(syn.x_mu, syn.y_mu)
        g.init_chains()
        g.run(1, a.iters)
        nb = 0
        for r in range(a.ranks):
            s = g.fetch_samples(r)
            li = g.fetch_likelihood(r)
            nb += s["hypo"].shape[0] * syn.n_events * 4 * (a.precision // 8)
        hist = g.get_histograms()
        cnt = g.get_counts()
        d2h = nb + hist.nbytes + 14 * 8
    torch.cuda.synchronize()
    e2e_wall = time.perf_counter() - te
    barrier()
    # ---- the only collective of the path: gather posterior histograms / reduce counters (NCCL) ----
    gathered = None
    if world > 1:
        from hypotremormcmc_b200.gather import gather_run
        tg = time.perf_counter()
        hist_all, counts_all = gather_run(g)
        torch.cuda.synchronize()
        gathered = {"histogram_events": int(hist_all.shape[0]), "hist_sum": int(hist_all.sum().item()),
                    "cold_proposals": int(counts_all[:7].sum().item()), "ms": (time.perf_counter() - tg) * 1e3}
    g.close()

    # ---- max over ranks ------------------------------------------------------------------------
    t = torch.tensor([dev_ms, e2e_wall, wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_wall_max, wall_max = [float(v) for v in t.tolist()]
    total_props = proposals_per_rank * world
    value = total_props / (dev_ms_max * 1e-3)
    e2e_value = e2e_n * a.iters * syn.n_events * a.ranks * a.chains * world / e2e_wall_max
    flop_per_prop = 30 * a.stations + 64
    achieved_tf = (proposals_per_rank / (dev_ms * 1e-3)) * flop_per_prop / 1e12

    if rank == 0:
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        # the ncu capture is of the default workload (BASELINE configs[1]); other shapes have no measured figure
        if os.path.exists(tp) and (a.events, a.stations, a.chains, a.ranks, a.precision) == (1000, 20, 16, 4, 32):
            try:
                traffic = json.load(open(tp)).get("fact_lane_kernel_dram_bytes_per_launch")
            except Exception:
                traffic = None
        # why the bound is not HBM: the kernel's measured DRAM traffic per launch against the measured copy bandwidth
        hbm = None
        mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if traffic and os.path.exists(mp):
            try:
                peak_gbs = float(json.load(open(mp))["hbm_gbs"])
                gbs = traffic / (dev_ms / a.steps * 1e-3) / 1e9
                hbm = {"achieved": gbs, "peak": peak_gbs, "unit": "GB/s", "frac": gbs / peak_gbs,
                       "peak_source": "MEASURED_PEAKS.json hbm_gbs"}
            except Exception:
                hbm = None
        cpu = None
        if not a.no_cpu:
            from oracle import pyoracle
            pyoracle.build()
            r, cores, sample = cpu_reference_rate(a, a.cpu_seconds)
            cpu = {"value": r, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": dev_ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f%d" % a.precision, "data": "synthetic",
            "config": {"workload": workload_name(a, world), "mode": "B (factorised, solve_* = F)",
                       "events_per_gpu": a.events, "stations": a.stations, "temperatures": a.chains,
                       "chains": a.ranks, "iterations_per_step": a.iters, "n_interval": a.interval,
                       "rng": "philox4x32-10", "kernel": {0: "auto", 1: "warp-per-chain", 2: "lane-per-chain"}[a.kernel],
                       "l2": "flushed between timed steps (256 MiB memset); state is register-resident per launch",
                       "cold_accept_rate": float(acc.sum() / max(1, p.sum()))},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_n, "what": "set_stations+set_observations+set_xy_prior (H2D) -> init_chains -> run -> "
                                            "fetch samples/likelihood/histograms/counts (D2H), host wall clock"},
            "gpu_launches": int(launches),
            "wall_ms_per_step_incl_flush": wall_max * 1e3 / a.steps,
            "clocks": clk,
            "roofline": {"bound": "fp32", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf, "traffic": traffic,
                         "peak_source": "own-measured FFMA microbenchmark (htm_measure_fp32_peak); "
                                        "MEASURED_PEAKS.json has no FP32 vector peak",
                         "flop_per_proposal": flop_per_prop, "mufu_gops_measured": mufu,
                         "kernel": "fact_lane_kernel" if a.kernel != 1 else "fact_warp_kernel", "hbm": hbm},
        }
        if gathered:
            line["nccl_gather"] = gathered
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
