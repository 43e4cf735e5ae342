#!/usr/bin/env python
"""bench.py -- MCMC proposals/s of the hypo_tremor_mcmc hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (chosen from BASELINE.json, named in config.workload):
  N = 1   configs[2]: 10,000 synthetic events x 50 stations x 16 temperatures x 4 chains (640,000 tempered
          chains), long chains with thinned sample output, factorised mode (solve_* = F), float32, Philox.
  N > 1   configs[3]: 100,000 events x 50 stations in total, 100,000/N per GPU (events shard, no data-path
          collective) -> "scaling": "strong"; the NCCL posterior-histogram gather + counter reduction is timed
          inside e2e and reported as e2e.collective_ms.
One "step" = `--iters` Metropolis iterations of every chain.  `--events` (per GPU) selects another size by hand.

Numbers on the JSON line:
  value     proposals/s, inputs resident in HBM, CUDA-event time of the K timed steps, max over ranks
  e2e       the same metric through the C ABI with HOST buffers: every step is a whole batch job
            (htm_set_stations/observations/xy_prior H2D -> htm_init_chains -> htm_run in chunks ->
            fetch samples + likelihood of every virtual rank, histograms, counts D2H [-> NCCL gather at N > 1]),
            host wall clock, max over ranks; e2e.breakdown_ms says where the time went
  roofline  achieved algorithmic FLOP/s ((30*S+64) per proposal) of the dominant kernel /
            own-measured FFMA peak (MEASURED_PEAKS.json holds no FP32 vector peak); roofline.traffic = DRAM bytes
            per launch of that kernel from the committed `ncu --set full` capture of this shape (null for shapes
            without a capture) and roofline.hbm = that traffic per launch time against MEASURED_PEAKS.json's copy
            bandwidth (the evidence that the kernel is not HBM-bound)
  extra     (N = 1) short device-resident runs of the other BASELINE shapes and instantiations on the same box:
            configs[1] (1000 x 20), the float64 instantiation, and mode C (blocked Gibbs, solve_* = T, the setting
            of sample/hypo_tremor.in) at the sample's 20 ranks x 5 chains; (N > 1) mode C with the events of every
            joint chain sharded over the GPUs (per-iteration exchange over NVLink peer memory)
  selfcheck (N > 1) a short event-sharded blocked-Gibbs run must equal the unsharded run of the same library
  cpu_baseline  the C++ restatement of the reference algorithm (oracle/, mode A: joint chain, one scalar per
            iteration, one swap per iteration), one thread per virtual rank, on a bounded sample of the same
            workload, with the rank x chain partition that runs fastest on this host.  The real Fortran/MPI binary
            cannot be built in this image (no gfortran/mpif90), so kind = "port".
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MCMC proposals/sec (events x chains)"
UNIT = "proposals/s"
SEED = 20231003  # 20231001 + config number


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--events", type=int, default=0,
                    help="events per GPU (0 = BASELINE: 10,000 at N = 1, 100,000/N at N > 1)")
    ap.add_argument("--stations", type=int, default=50)
    ap.add_argument("--ranks", type=int, default=4, help="virtual ranks = independent tempering groups per event")
    ap.add_argument("--chains", type=int, default=16, help="chains (temperatures) per rank")
    ap.add_argument("--iters", type=int, default=20000, help="iterations per step")
    ap.add_argument("--interval", type=int, default=1000)
    ap.add_argument("--chunks", type=int, default=4, help="e2e: htm_run calls per step (samples drain between them)")
    ap.add_argument("--precision", type=int, default=32, choices=[32, 64])
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 warp-per-chain, 2 lane-per-chain")
    ap.add_argument("--slots", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target duration of the CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    return ap.parse_args()


def events_total(a, world):
    if a.events > 0:
        return a.events * world, "weak"
    return (10000, "strong") if world == 1 else (100000, "strong")


def config_block(a, world):
    """Identical in both arms (the driver compares it key for key)."""
    E_total, _ = events_total(a, world)
    shape = (E_total, a.stations, a.chains, a.ranks)
    if shape == (10000, 50, 16, 4) and world == 1:
        tag = "BASELINE configs[2]"
    elif shape == (100000, 50, 16, 4) and world > 1:
        tag = "BASELINE configs[3]"
    elif shape == (1000, 20, 16, 4):
        tag = "BASELINE configs[1]"
    else:
        tag = "not a BASELINE config"
    return {
        "workload": "%s: %d synthetic events x %d stations x %d temperatures x %d chains, thinned sample output, "
                    "on %d GPU%s" % (tag, E_total, a.stations, a.chains, a.ranks, world, "" if world == 1 else "s"),
        "events": E_total, "events_per_gpu": E_total // world, "stations": a.stations, "temperatures": a.chains,
        "chains": a.ranks, "n_interval": a.interval, "solve": "F", "seed": SEED,
        "l2": "GPU arm: flushed between timed steps (256 MiB memset); chain state is register-resident per launch",
    }


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
class CpuReference:
    """The reference algorithm (joint chains, src/hypo_tremor_mcmc.f90:236-284) on the host cores: one thread per
    virtual rank, the total number of chains of the GPU arm, split into ranks x chains the way that runs fastest
    here (ranks meet at a barrier once per iteration for the swap, so more threads is not always faster)."""

    def __init__(self, n_events, n_sta, total_chains, solve, seed=SEED, probe_seconds=0.7, n0=500):
        import hypotremormcmc_b200 as H
        from oracle.pyoracle import Oracle
        self.H, self.Oracle = H, Oracle
        self.E, self.S, self.total, self.solve = n_events, n_sta, total_chains, solve
        self.syn = H.Synthetic(n_events, n_sta, seed)
        cores = os.cpu_count() or 1
        # candidate partitions: divisors of the chain total that fit the cores and keep >= 2 chains per rank
        cand = [d for d in range(1, total_chains + 1) if total_chains % d == 0 and d <= cores and total_chains // d >= 2]
        cand = sorted(set(cand[-4:]))  # the four largest are the only plausible winners
        if probe_seconds <= 0:
            cand = cand[-1:]
        self.n0 = n0
        best = None
        self.probes = {}
        for ranks in cand:  # short probe of each partition: iterations done / wall
            o = self._make(ranks)
            done, wall, n, it = 0, 0.0, max(1, self.n0 // 2), 3
            while wall < probe_seconds or done == 0:
                wall += o.run_threaded(it, it + n - 1)
                done += n
                it += n
                n *= 2
            self.probes[ranks] = done * total_chains / wall
            o.close()
            if best is None or self.probes[ranks] > self.probes[best]:
                best = ranks
        self.ranks, self.chains = best, total_chains // best
        self.o = self._make(best)
        self.it = 3

    def _make(self, ranks):
        H = self.H
        s = 1 if self.solve else 0
        cfg = H.default_config(n_sta=self.S, n_events=self.E, n_procs=ranks, n_chains=self.total // ranks, n_cool=1,
                               n_iter=2 ** 31 - 2, n_burn=2 ** 31 - 2, n_interval=1000, mode=H.MODE_REPLAY, precision=64,
                               solve_vs=s, solve_t_corr=s, solve_qs=s, solve_a_corr=s)
        o = self.Oracle(cfg, self.syn)
        o.init_chains()
        o.run_threaded(1, 2)  # iteration 1 is a full O(E*S) likelihood per chain: outside every timed sample
        return o

    def sample(self, seconds):
        """proposals/s over about `seconds` of wall clock, continuing the chains"""
        done, wall, n = 0, 0.0, self.n0
        while wall < seconds:
            t = self.o.run_threaded(self.it, self.it + n - 1)
            wall += t
            done += n
            self.it += n
            if t < seconds / 8:
                n *= 2
        self.last = (done, wall)
        return done * self.total / wall

    def describe(self):
        done, wall = self.last
        return ("oracle mode A (reference semantics, solve_* = %s), %d events x %d stations, %d virtual ranks x %d chains "
                "(fastest of the partitions %s), %d iterations in %.1f s, %d threads"
                % ("T" if self.solve else "F", self.E, self.S, self.ranks, self.chains,
                   {k: "%.3g/s" % v for k, v in self.probes.items()}, done, wall, self.ranks))

    def close(self):
        self.o.close()


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle
    pyoracle.build()
    world = max(1, a.gpus)
    E_total, scaling = events_total(a, world)
    # each step a bounded sample; the whole --steps K --warmup W run stays within ~2 minutes
    per_step = min(a.cpu_seconds, max(1.0, min(20.0, 100.0 / max(1, a.steps + a.warmup))))
    ref = CpuReference(E_total, a.stations, a.ranks * a.chains, solve=False)
    rates = []
    for i in range(a.warmup + a.steps):
        r = ref.sample(per_step)
        if i >= a.warmup:
            rates.append(r)
    v = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_block(a, world),
        "details": {"mode": "A (reference joint chain)",
                    "note": "C++ restatement of the reference algorithm; the Fortran/MPI binary cannot be built here"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": ref.ranks, "kind": "port", "sample": ref.describe()},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    ref.close()
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
class DeviceTimer:
    """K timed htm_run steps with an L2 flush before each; CUDA-event time from the library's own stream."""

    def __init__(self, torch):
        self.torch = torch
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def run(self, g, iters, warmup, steps, it0=1):
        torch = self.torch
        for _ in range(warmup):
            g.run(it0, it0 + iters - 1)
            g.synchronize()
            g.discard_samples()
            it0 += iters
        ms_steps, launches = [], 0
        for _ in range(steps):
            self.flush.zero_()  # outside the CUDA-event pair
            torch.cuda.synchronize()
            g.run(it0, it0 + iters - 1)
            g.synchronize()
            ms, nl, _ = g.last_run_stats()
            ms_steps.append(ms)
            launches += nl
            g.discard_samples()
            it0 += iters
        return ms_steps, launches, it0


def traffic_of(kernel, shape):
    """DRAM bytes per launch from the committed ncu capture of exactly this shape, else None."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(tp))
    except Exception:
        return None, None
    for ent in t.get("captures", []):
        if ent.get("kernel") == kernel and tuple(ent.get("shape", ())) == tuple(shape):
            return ent.get("dram_bytes_per_launch"), ent.get("source")
    return None, None


def hbm_block(traffic, ms_per_launch):
    mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if not traffic or not os.path.exists(mp):
        return None
    try:
        peak_gbs = float(json.load(open(mp))["hbm_gbs"])
    except Exception:
        return None
    gbs = traffic / (ms_per_launch * 1e-3) / 1e9
    return {"achieved": gbs, "peak": peak_gbs, "unit": "GB/s", "frac": gbs / peak_gbs,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs"}


def extra_mode_b(H, timer, a, local, E, S, K, R, precision, iters, peak_tf, peak64_tf=None):
    """short device-resident run of another mode-B shape / instantiation"""
    syn = H.Synthetic(E, S, SEED - 1)
    cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=iters, n_burn=0,
                           n_interval=a.interval, mode=H.MODE_FACTORISED, solve_vs=0, solve_t_corr=0, solve_qs=0,
                           solve_a_corr=0, precision=precision, device=local, max_samples=2 * (iters // a.interval + 2),
                           hist_bins=64, seed=SEED - 1)
    with H.HypoTremorB200(cfg) as g:
        g.load(syn)
        g.init_chains()
        ms, nl, _ = timer.run(g, iters, 3, 5)
    props = 5 * iters * E * R * K
    rate = props / (sum(ms) * 1e-3)
    flop = 30 * S + 64
    tr, src = traffic_of("fact_lane_kernel", (E, S, K, R, precision))
    return {"workload": "%d events x %d stations x %d temperatures x %d chains, mode B, f%d" % (E, S, K, R, precision),
            "value": rate, "unit": UNIT, "ms_per_step": sum(ms) / 5, "iterations_per_step": iters, "steps": 5, "warmup": 3,
            "gpu_launches": nl,
            "roofline": {"bound": "fp%d" % precision, "achieved": rate * flop / 1e12,
                         "peak": peak_tf if precision == 32 else peak64_tf, "unit": "TFLOP/s",
                         "frac": rate * flop / 1e12 / (peak_tf if precision == 32 else peak64_tf),
                         "flop_per_proposal": flop, "traffic": tr, "traffic_source": src,
                         "note": None if precision == 32 else
                         "peak = own-measured DFMA microbenchmark (htm_measure_fp64_peak); the float64 kernel keeps the "
                         "reference's operation order (IEEE sqrt, division, libdevice log), so most of its FP64 work is "
                         "inside those, not in the 30 S algorithmic operations"}}


def extra_mode_c(H, timer, a, local, E, S, R, K, iters, peak_tf, cpu_seconds):
    """blocked Gibbs at the sample file's chain layout, solve_* = T; one proposal = one (chain, event) hypocentre
    step (which also evaluates the chain's pending shared-parameter proposal for that event: 2 x 30 S FLOP)"""
    syn = H.Synthetic(E, S, SEED + 10)
    cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=10 ** 6, n_burn=0,
                           n_interval=50, mode=H.MODE_BLOCKED_GIBBS, precision=32, device=local, max_samples=0,
                           seed=SEED + 10)
    with H.HypoTremorB200(cfg) as g:
        g.load(syn)
        g.init_chains()
        ms, nl, _ = timer.run(g, iters, 2, 5)
        p, acc = g.get_counts()
    J = R * K
    props = 5 * iters * (E + 1) * J
    rate = props / (sum(ms) * 1e-3)
    flop = 2 * 30 * S + 64
    out = {"workload": "%d events x %d stations, %d ranks x %d chains = %d joint chains, solve_* = T (sample/hypo_tremor.in), "
                       "mode C blocked Gibbs, f32" % (E, S, R, K, J),
           "value": rate, "unit": UNIT, "us_per_iteration": sum(ms) * 1e3 / (5 * iters), "iterations_per_step": iters,
           "steps": 5, "warmup": 2, "gpu_launches": nl,
           "cold_accept_rate_hypo": float(acc[4:].sum() / max(1, p[4:].sum())),
           "cold_accept_rate_shared": float(acc[:4].sum() / max(1, p[:4].sum())),
           "roofline": {"bound": "fp32", "achieved": rate * flop / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                        "frac": rate * flop / 1e12 / peak_tf, "flop_per_proposal": flop,
                        "kernel": "gibbs_f32_kernel", "executed_flop_per_proposal": 50 * S + 200,
                        "note": "algorithmic count = two 30 S evaluations as the reference does; the kernel evaluates one "
                                "pass with moments (25 packed operations per station pair) + an O(1) delta"}}
    if cpu_seconds > 0:
        ref = CpuReference(E, S, J, solve=True, seed=SEED + 10, probe_seconds=0, n0=20)
        r = ref.sample(cpu_seconds)
        out["cpu_baseline"] = {"value": r, "unit": UNIT, "cores": ref.ranks, "kind": "port", "sample": ref.describe()}
        ref.close()
    return out


def extra_upstream(H, a, local, peak64_tf):
    """the two upstream stages built beside the hot path, each timed by its own CUDA events
    inside the call, with the oracle timed on a bounded sample as the CPU baseline"""
    import time
    from hypotremormcmc_b200 import api
    out = {}
    E, S = 100000, 50
    syn = H.Synthetic(E, S, SEED + 20)
    args = (syn.sta_x, syn.sta_y, syn.sta_z, 7.0, syn.t_obs, syn.t_stdv, syn.a_obs, syn.a_stdv)
    ms = min(api.select_events(*args, device=local)["kernel_ms"] for _ in range(4))
    nbytes = E * (32 * S + 52)
    hb = hbm_block(nbytes, ms)
    out["select"] = {"workload": "hypo_tremor_select regression + acceptance, %d windows x %d stations, f64" % (E, S),
                     "value": E / (ms * 1e-3), "unit": "windows/s", "kernel_ms": ms, "gpu_launches": 1,
                     "roofline": {"bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": hb["peak"] if hb else None,
                                  "unit": "GB/s", "frac": hb["frac"] if hb else None, "traffic": None,
                                  "bytes_per_window": 32 * S + 52, "kernel": "select_kernel",
                                  "note": "FP64 sqrt / log / divisions per station keep it far from the HBM bound"}}
    W, n = 2000, 300
    rng = np.random.default_rng(SEED)
    kern = np.hanning(21)
    env = np.stack([np.convolve(rng.normal(0, 1, (n // 2) * (W + 1)) ** 2, kern, mode="same") for _ in range(S)])
    win = np.arange(1, W + 1)
    ms = min(api.measure_windows(env, 1.0, n, n // 2, win, device=local)["kernel_ms"] for _ in range(3))
    flop = 2.0 * W * (S * (S - 1) // 2) * n * n
    out["measure"] = {"workload": "hypo_tremor_measure optimize_cc + optimize_amp, %d windows x %d stations x %d samples, f64"
                                  % (W, S, n),
                      "value": W / (ms * 1e-3), "unit": "windows/s", "kernel_ms": ms, "gpu_launches": 1,
                      "roofline": {"bound": "fp64", "achieved": flop / (ms * 1e-3) / 1e12, "peak": peak64_tf,
                                   "unit": "TFLOP/s", "frac": flop / (ms * 1e-3) / 1e12 / peak64_tf, "traffic": None,
                                   "flop_per_window": flop / W, "kernel": "measure_kernel",
                                   "peak_source": "own-measured DFMA microbenchmark (htm_measure_fp64_peak)"}}
    day = np.stack([np.convolve(rng.normal(0, 1, 86400) ** 2, kern, mode="same") for _ in range(S)])
    ms = min(api.detect_windows(day, n, n // 2, 0.98, 300, device=local)["kernel_ms"] for _ in range(3))
    n_win = (86400 - n) // (n // 2)
    flop = 2.0 * n_win * (S * (S - 1) // 2) * n * n
    out["detect"] = {"workload": "hypo_tremor_measure scan_cc with the correlation functions of hypo_tremor_correlate recomputed: "
                                 "one day of %d stations at 1 sample/s = %d windows x %d pairs x %d lags, f64"
                                 % (S, n_win, S * (S - 1) // 2, n),
                     "value": n_win / (ms * 1e-3), "unit": "windows/s", "kernel_ms": ms, "gpu_launches": 3,
                     "roofline": {"bound": "fp64", "achieved": flop / (ms * 1e-3) / 1e12, "peak": peak64_tf, "unit": "TFLOP/s",
                                  "frac": flop / (ms * 1e-3) / 1e12 / peak64_tf, "traffic": None,
                                  "kernel": "measure_kernel (mode 1) + quantile_select_kernel + detect_kernel",
                                  "note": "the time includes 3-4 selection scans of the %.1f GB of correlation values"
                                          % (8.0 * n_win * n * (S * (S - 1) // 2) / 1e9)}}
    if not a.no_cpu:
        from oracle import pyoracle
        t0 = time.time()
        pyoracle.detect_windows(day[:, :n + 16 * (n // 2)], n, n // 2, 0.98, 300)
        out["detect"]["cpu_baseline"] = {"value": 16 / (time.time() - t0), "unit": "windows/s", "cores": 1, "kind": "port",
                                         "sample": "oracle detect_windows (direct float64 sums, full sort) on the first 16 windows"}
        t0 = time.time()
        pyoracle.select_events(*[v[:20000] if getattr(v, "ndim", 0) == 2 else v for v in args])
        out["select"]["cpu_baseline"] = {"value": 20000 / (time.time() - t0), "unit": "windows/s", "cores": 1, "kind": "port",
                                         "sample": "oracle/htm_oracle_select.hpp on the first 20 000 windows"}
        t0 = time.time()
        pyoracle.measure_windows(env, 1.0, n, n // 2, win[:16])
        out["measure"]["cpu_baseline"] = {"value": 16 / (time.time() - t0), "unit": "windows/s", "cores": 1, "kind": "port",
                                          "sample": "oracle/htm_oracle_measure.hpp (direct float64 sums) on the first 16 windows"}
    return out


def p2p_setup(g, dist, world, rank):
    ids = [g.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    g.comm_init(ids[0])
    handles = [None] * world
    dist.all_gather_object(handles, g.comm_p2p_export())
    g.comm_p2p_import(handles)


def multi_gpu_mode_c(H, dist, torch, a, world, rank, local, peak_tf):
    """(N > 1) selfcheck: event-sharded blocked Gibbs == unsharded run; and the throughput of ONE joint ensemble
    whose events are sharded over the GPUs (the one real exchange step of the path, fused into the sweep)."""
    ok = True
    # ---- selfcheck (small, traced) ----
    E, S, R, K, n_it = 2011, 20, 2, 4, 60
    syn = H.Synthetic(E, S, 9)
    base = dict(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=n_it, n_burn=10, n_interval=5,
                mode=H.MODE_BLOCKED_GIBBS, precision=32, max_samples=32)
    sh = syn.shard(rank, world)
    with H.HypoTremorB200(H.default_config(device=local, shard_rank=rank, shard_count=world, gibbs_shard_events=1, **base)) as g:
        g.load(sh)
        g.init_chains()
        p2p_setup(g, dist, world, rank)
        tr, sw = g.run_traced(1, 40)
        g.run(41, n_it)
        st = g.get_chain_state(1, 2)
        _, p, acc = g.gather(histograms=False)
    with H.HypoTremorB200(H.default_config(device=local, **base)) as u:
        u.load(syn)
        u.init_chains()
        tr_u, sw_u = u.run_traced(1, 40)
        u.run(41, n_it)
        su = u.get_chain_state(1, 2)
        pu, au = u.get_counts()
    lo = sh.event_offset
    for f in ("proposal_type", "prior_ok", "accepted"):
        ok &= bool(np.array_equal(tr[f][:, :-1], tr_u[f][:, lo:lo + sh.n_events]) and np.array_equal(tr[f][:, -1], tr_u[f][:, -1]))
    ok &= bool(np.array_equal(sw, sw_u))
    # (the float64 sums of float32 terms depend on the grouping into partial sums in their last bits only)
    ok &= bool(np.all(np.abs(tr["log_likelihood"][:, -1] - tr_u["log_likelihood"][:, -1]) <= 1e-13 * np.abs(tr_u["log_likelihood"][:, -1])))
    ok &= st["vs"] == su["vs"] and st["qs"] == su["qs"] and st["temp"] == su["temp"]
    ok &= bool(np.array_equal(st["hypo"], su["hypo"][3 * lo:3 * (lo + sh.n_events)]))
    ok &= bool(np.array_equal(p, pu) and np.array_equal(acc, au))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    selfcheck = {"event_sharded_gibbs_equals_unsharded": bool(int(flag.item()) == 1),
                 "what": "%d events x %d stations, %d joint chains, f32, %d iterations traced + %d more: every flag, "
                         "swap, counter and final state identical on every rank, summed log-likelihoods equal to 1e-13 "
                         "(peer-memory exchange inside the persistent sweep)" % (E, S, R * K, 40, n_it - 40)}
    # ---- throughput: 100,000 x 50, 20 joint chains, events sharded ----
    E, S, R, K, n_it = 100000, 50, 4, 5, 300
    syn = H.Synthetic(E, S, SEED + 11).shard(rank, world)
    cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=10 ** 6, n_burn=0, n_interval=50,
                           mode=H.MODE_BLOCKED_GIBBS, precision=32, device=local, shard_rank=rank, shard_count=world,
                           gibbs_shard_events=1, seed=SEED + 11)
    with H.HypoTremorB200(cfg) as g:
        g.load(syn)
        g.init_chains()
        p2p_setup(g, dist, world, rank)
        g.run(1, 20)
        g.synchronize()
        dist.barrier()
        g.run(21, 20 + n_it)
        g.synchronize()
        ms, nl, _ = g.last_run_stats()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    rate = n_it * (E + 1) * R * K / (ms * 1e-3)
    flop = 2 * 30 * S + 64
    perf = {"workload": "ONE ensemble of %d joint chains over %d events x %d stations, solve_* = T, events sharded over "
                        "%d GPUs, per-iteration sum exchanged over NVLink peer memory inside the sweep kernel" % (R * K, E, S, world),
            "value": rate, "unit": UNIT, "us_per_iteration": ms * 1e3 / n_it, "iterations": n_it, "scaling": "strong",
            "roofline": {"bound": "fp32", "achieved": rate * flop / 1e12, "peak": peak_tf * world, "unit": "TFLOP/s",
                         "frac": rate * flop / 1e12 / (peak_tf * world), "flop_per_proposal": flop}}
    return selfcheck, perf


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
    E_total, scaling = events_total(a, world)
    syn_all = H.Synthetic(E_total, a.stations, SEED)
    syn = syn_all.shard(rank, world)
    del syn_all
    n_rec = (a.iters + a.interval - 1) // a.interval + 1
    hist_bins = 64
    cfg = H.default_config(
        n_sta=a.stations, n_events=E_total, n_procs=a.ranks, n_chains=a.chains, n_cool=1,
        n_iter=a.iters, n_burn=0, n_interval=a.interval, mode=H.MODE_FACTORISED,
        solve_vs=0, solve_t_corr=0, solve_qs=0, solve_a_corr=0, precision=a.precision,
        ladder=H.LADDER_RANDOM, kernel=a.kernel, lane_slots=a.slots, device=local,
        shard_rank=rank, shard_count=world, max_samples=n_rec, hist_bins=hist_bins, seed=SEED)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    timer = DeviceTimer(torch)

    # ---- device-resident arm: value ----------------------------------------------------------
    g = H.HypoTremorB200(cfg)
    g.load(syn)
    g.init_chains()
    g.synchronize()
    peak_tf, mufu = measure_fp32_peak(local)
    _, _, it0 = timer.run(g, a.iters, a.warmup, 0)
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    t0 = time.perf_counter()
    ms_steps, launches, it0 = timer.run(g, a.iters, 0, a.steps, it0)
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
    n_chunks = max(1, min(a.chunks, a.iters // max(1, a.interval)))
    bounds = [1 + (a.iters * c) // n_chunks for c in range(n_chunks + 1)]
    bd = {"inputs": 0.0, "init": 0.0, "run_and_fetch": 0.0, "hist_counts": 0.0, "collective": 0.0}
    if world > 1:
        from hypotremormcmc_b200.gather import gather_run
    for i in range(2 + e2e_n):
        if i == 2:
            barrier()
            bd = dict.fromkeys(bd, 0.0)
            te = time.perf_counter()
        t1 = time.perf_counter()
        g.set_stations(syn.sta_x, syn.sta_y, syn.sta_z)
        g.set_observations(syn.t_obs, syn.t_stdv, syn.a_obs, syn.a_stdv)
        g.set_xy_prior(syn.x_mu, syn.y_mu)
        t2 = time.perf_counter()
        g.init_chains()
        t3 = time.perf_counter()
        for c in range(n_chunks):  # the driver's loop: run a chunk, drain what the previous chunks recorded
            g.run(bounds[c], bounds[c + 1] - 1)
        nb = 0
        for r in range(a.ranks):
            s = g.fetch_samples(r)
            li = g.fetch_likelihood(r)
            nb += s["hypo"].shape[0] * syn.n_events * 4 * (a.precision // 8)
        t4 = time.perf_counter()
        hist = g.get_histograms()
        cnt = g.get_counts()
        t5 = time.perf_counter()
        d2h = nb + hist.nbytes + 14 * 8
        if world > 1:  # the only collective of the path: posterior histograms all-gather + counter reduction (NCCL)
            hist_all, counts_all = gather_run(g)
            torch.cuda.synchronize()
        t6 = time.perf_counter()
        for k, v in zip(bd, (t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5)):
            bd[k] += v * 1e3
    torch.cuda.synchronize()
    e2e_wall = time.perf_counter() - te
    barrier()
    gathered = None
    if world > 1:
        gathered = {"histogram_events": int(hist_all.shape[0]), "hist_sum": int(hist_all.sum().item()),
                    "cold_proposals": int(counts_all[:7].sum().item())}
    g.close()

    # ---- max over ranks ------------------------------------------------------------------------
    t = torch.tensor([dev_ms, e2e_wall, wall, bd["collective"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_wall_max, wall_max, coll_ms_max = [float(v) for v in t.tolist()]
    total_props = proposals_per_rank * world  # shards differ by at most one event; rank 0 holds the largest
    value = total_props / (dev_ms_max * 1e-3)
    e2e_value = e2e_n * a.iters * syn.n_events * a.ranks * a.chains * world / e2e_wall_max
    flop_per_prop = 30 * a.stations + 64
    achieved_tf = (proposals_per_rank / (dev_ms * 1e-3)) * flop_per_prop / 1e12

    # ---- other shapes / modes on the same box ----------------------------------------------------
    extra, selfcheck = {}, None
    if not a.no_extra:
        if world == 1:
            extra["configs1_f32"] = extra_mode_b(H, timer, a, local, 1000, 20, 16, 4, 32, 20000, peak_tf)
            from hypotremormcmc_b200.api import measure_fp64_peak
            extra["configs2_f64"] = extra_mode_b(H, timer, a, local, a.events or 10000, a.stations, a.chains, a.ranks, 64,
                                                 500, peak_tf, measure_fp64_peak(local))
            if not a.no_cpu:
                from oracle import pyoracle
                pyoracle.build()
            extra["mode_c_10k"] = extra_mode_c(H, timer, a, local, 10000, 50, 20, 5, 100, peak_tf,
                                               0 if a.no_cpu else min(6.0, a.cpu_seconds))
            extra["mode_c_100k"] = extra_mode_c(H, timer, a, local, 100000, 50, 20, 5, 20, peak_tf, 0)
            extra["upstream"] = extra_upstream(H, a, local, measure_fp64_peak(local))
        else:
            selfcheck, extra["mode_c_event_sharded"] = multi_gpu_mode_c(H, dist, torch, a, world, rank, local, peak_tf)

    if rank == 0:
        kname = "fact_lane_kernel" if a.kernel != 1 else "fact_warp_kernel"
        traffic, tsrc = traffic_of(kname, (syn.n_events, a.stations, a.chains, a.ranks, a.precision))
        cpu = None
        if not a.no_cpu and world == 1:
            from oracle import pyoracle
            pyoracle.build()
            ref = CpuReference(E_total, a.stations, a.ranks * a.chains, solve=False)
            r = ref.sample(a.cpu_seconds)
            cpu = {"value": r, "unit": UNIT, "cores": ref.ranks, "kind": "port", "sample": ref.describe()}
            ref.close()
        e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "steps": e2e_n, "ratio_to_value": e2e_value / value,
               "breakdown_ms_per_step": {k: v / e2e_n for k, v in bd.items()},
               "what": "set_stations+set_observations+set_xy_prior (H2D) -> init_chains -> run (%d chunks) -> fetch "
                       "samples/likelihood of every rank, histograms, counts (D2H)%s, host wall clock"
                       % (n_chunks, " -> NCCL histogram all-gather + counter all-reduce" if world > 1 else "")}
        if world > 1:
            e2e["collective_ms"] = coll_ms_max / e2e_n
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": dev_ms_max / a.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f%d" % a.precision, "data": "synthetic",
            "config": config_block(a, world),
            "details": {"mode": "B (factorised, solve_* = F)", "iterations_per_step": a.iters, "rng": "philox4x32-10",
                        "kernel": {0: "auto", 1: "warp-per-chain", 2: "lane-per-chain"}[a.kernel],
                        "cold_accept_rate": float(acc.sum() / max(1, p.sum()))},
            "e2e": e2e,
            "gpu_launches": int(launches),
            "wall_ms_per_step_incl_flush": wall_max * 1e3 / a.steps,
            "clocks": clk,
            "roofline": {"bound": "fp32", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf, "traffic": traffic, "traffic_source": tsrc,
                         "peak_source": "own-measured FFMA microbenchmark (htm_measure_fp32_peak); "
                                        "MEASURED_PEAKS.json has no FP32 vector peak",
                         "flop_per_proposal": flop_per_prop, "mufu_gops_measured": mufu,
                         "mufu_cap_proposals_per_s": mufu * 1e9 / (2 * a.stations + 6),
                         "kernel": kname, "hbm": hbm_block(traffic, dev_ms / a.steps)},
        }
        if gathered:
            line["nccl_gather"] = gathered
        if selfcheck:
            line["selfcheck"] = selfcheck
        if extra:
            line["extra"] = extra
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
