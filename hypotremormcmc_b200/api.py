"""Python host-side mirror of the C ABI (include/htm_b200.h), via ctypes.

The calls map one-to-one onto the exported functions, which in turn replace the Fortran
type-bound calls of the reference driver (src/hypo_tremor_mcmc.f90:101-118,188-209,
236-291).  Nothing here computes: without the CUDA library (or without a GPU) every
compute call raises.
"""
import ctypes
import os

import numpy as np

from .config import (HtmConfig, StepTrace, SwapTrace, HTM_OK, MODE_FACTORISED, MODE_REPLAY, MODE_BLOCKED_GIBBS,
                     STEP_TRACE_DTYPE, SWAP_TRACE_DTYPE, copy_config)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

EXPORTS = [
    "htm_config_default", "htm_create", "htm_destroy", "htm_last_error", "htm_set_stations",
    "htm_set_observations", "htm_set_xy_prior", "htm_set_globals", "htm_init_chains",
    "htm_set_chain_state", "htm_get_chain_state", "htm_loglik", "htm_run", "htm_run_traced",
    "htm_synchronize", "htm_replay", "htm_fetch_samples", "htm_fetch_likelihood",
    "htm_discard_samples", "htm_get_counts", "htm_get_histograms", "htm_device_ptr",
    "htm_last_run_stats", "htm_measure_fp32_peak", "htm_comm_unique_id", "htm_comm_init", "htm_gather",
    "htm_comm_p2p_export", "htm_comm_p2p_import", "htm_gather_samples", "htm_gibbs_pending", "htm_gibbs_last_sums",
    "htm_posterior_quantiles", "htm_measure_fp64_peak", "htm_select_events", "htm_measure_windows", "htm_detect_windows",
]


class HtmError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libhtm_b200 error %d: %s" % (code, message))
        self.code = code
        self.message = message


def library_path():
    # HTM_B200_LIB: load another build of the same library (kernel-tuning experiments)
    return os.environ.get("HTM_B200_LIB") or os.path.join(_HERE, "csrc", "libhtm_b200.so")


def load_library():
    """Load csrc/libhtm_b200.so.  Fails loudly when it is missing: there is no fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(
            "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C hypotremormcmc_b200/csrc` (there is no CPU fallback)" % path)
    lib = ctypes.CDLL(path)
    dp = ctypes.POINTER(ctypes.c_double)
    ip = ctypes.POINTER(ctypes.c_int32)
    lp = ctypes.POINTER(ctypes.c_int64)
    vp = ctypes.c_void_p
    i32 = ctypes.c_int32
    sig = {
        "htm_config_default": [ctypes.POINTER(HtmConfig)],
        "htm_create": [ctypes.POINTER(vp), ctypes.POINTER(HtmConfig)],
        "htm_destroy": [vp],
        "htm_last_error": [vp, ctypes.c_char_p, i32],
        "htm_set_stations": [vp, dp, dp, dp],
        "htm_set_observations": [vp, dp, dp, dp, dp],
        "htm_set_xy_prior": [vp, dp, dp],
        "htm_set_globals": [vp, ctypes.c_double, ctypes.c_double, dp, dp],
        "htm_init_chains": [vp],
        "htm_set_chain_state": [vp, i32, i32, dp, dp, dp, ctypes.c_double, ctypes.c_double,
                                ctypes.c_double, ctypes.c_double],
        "htm_get_chain_state": [vp, i32, i32, dp, dp, dp, dp, dp, dp, dp],
        "htm_loglik": [vp, i32, dp, dp, dp, dp, dp, dp, dp],
        "htm_run": [vp, i32, i32],
        "htm_run_traced": [vp, i32, i32, vp, vp],
        "htm_synchronize": [vp],
        "htm_replay": [vp, i32, i32, ctypes.POINTER(ip), lp, vp, vp, lp],
        "htm_fetch_samples": [vp, i32, i32, ip, ip, dp, dp, dp, dp, dp],
        "htm_fetch_likelihood": [vp, i32, i32, ip, ip, dp],
        "htm_discard_samples": [vp],
        "htm_get_counts": [vp, lp, lp],
        "htm_get_histograms": [vp, ctypes.POINTER(ctypes.c_uint32)],
        "htm_device_ptr": [vp, i32, ctypes.POINTER(vp), lp],
        "htm_last_run_stats": [vp, dp, lp, lp],
        "htm_measure_fp32_peak": [i32, dp, dp],
        "htm_measure_fp64_peak": [i32, dp],
        "htm_select_events": [i32, i32, i32, dp, dp, dp, ctypes.c_double, dp, dp, dp, dp, ctypes.c_double, ctypes.c_double,
                              ctypes.c_double, ctypes.c_double, dp, dp, dp, dp, dp, dp, ip, dp],
        "htm_measure_windows": [i32, i32, ctypes.c_int64, dp, ctypes.c_double, i32, i32, i32, ip, dp, dp, dp, dp, ip, dp],
        "htm_detect_windows": [i32, i32, ctypes.c_int64, dp, i32, i32, ctypes.c_double, i32, i32, dp, dp, ip, ip, dp],
        "htm_comm_unique_id": [ctypes.c_char_p],
        "htm_comm_init": [vp, ctypes.c_char_p],
        "htm_gather": [vp, ctypes.POINTER(ctypes.c_uint32), lp, lp],
        "htm_gather_samples": [vp, i32, i32, ctypes.POINTER(i32), ctypes.POINTER(i32), dp, dp, dp, dp, dp],
        "htm_posterior_quantiles": [vp, ip, dp, dp, dp, dp, dp],
        "htm_gibbs_pending": [vp, ip, ip, dp],
        "htm_gibbs_last_sums": [vp, dp, dp],
        "htm_comm_p2p_export": [vp, ctypes.c_char_p],
        "htm_comm_p2p_import": [vp, ctypes.c_char_p],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = ctypes.c_int32
    _LIB = lib
    return lib


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if a is not None else None


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError("expected shape %r, got %r" % (shape, a.shape))
    return a


def measure_fp32_peak(device=0):
    lib = load_library()
    tf, mu = ctypes.c_double(), ctypes.c_double()
    rc = lib.htm_measure_fp32_peak(device, ctypes.byref(tf), ctypes.byref(mu))
    if rc != HTM_OK:
        raise HtmError(rc, "htm_measure_fp32_peak failed (no CUDA device?)")
    return tf.value, mu.value


def select_events(sta_x, sta_y, sta_z, z_guess, t, t_err, a, a_err, vs_min=2.0, vs_max=4.0, b_min=0.015, b_max=0.03,
                  device=0):
    """hypo_tremor_select for all windows at once (defaults: sample/hypo_tremor.in:109-117).  t, t_err, a, a_err:
    [n_events, n_sta].  Returns dict(vs, t0, b, a0, cc_t, cc_a, selected, kernel_ms)."""
    lib = load_library()
    t = _f64(t)
    E, S = t.shape
    arr = [_f64(v, (S,)) for v in (sta_x, sta_y, sta_z)] + [t] + [_f64(v, (E, S)) for v in (t_err, a, a_err)]
    out = [np.empty(E) for _ in range(6)]
    sel = np.zeros(E, dtype=np.int32)
    ms = ctypes.c_double()
    rc = lib.htm_select_events(device, S, E, _dptr(arr[0]), _dptr(arr[1]), _dptr(arr[2]), z_guess, _dptr(arr[3]),
                               _dptr(arr[4]), _dptr(arr[5]), _dptr(arr[6]), vs_min, vs_max, b_min, b_max,
                               *[_dptr(o) for o in out], sel.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ctypes.byref(ms))
    if rc != HTM_OK:
        buf = ctypes.create_string_buffer(512)
        lib.htm_last_error(None, buf, 512)
        raise HtmError(rc, buf.value.decode())
    return dict(vs=out[0], t0=out[1], b=out[2], a0=out[3], cc_t=out[4], cc_a=out[5], selected=sel, kernel_ms=ms.value)


def measure_windows(env, dt, n_smp, n_step, win_id, want_lag=False, device=0):
    """hypo_tremor_measure's lag / amplitude optimisation for all detected windows at once.  env: [n_sta, n_total] merged
    envelopes sampled every dt; window w = samples (win_id[w] - 1) * n_step ... + n_smp - 1.  Returns dict(t, t_stdv, amp,
    amp_stdv [n_win, n_sta], kernel_ms and, if asked for, lag [n_win, n_sta (n_sta - 1) / 2])."""
    lib = load_library()
    env = _f64(env)
    S, n_total = env.shape
    win_id = np.ascontiguousarray(win_id, dtype=np.int32)
    W = win_id.size
    out = [np.empty((W, S)) for _ in range(4)]
    lag = np.zeros((W, S * (S - 1) // 2), dtype=np.int32) if want_lag else None
    ms = ctypes.c_double()
    i32p = ctypes.POINTER(ctypes.c_int32)
    rc = lib.htm_measure_windows(device, S, n_total, _dptr(env), dt, n_smp, n_step, W, win_id.ctypes.data_as(i32p),
                                 *[_dptr(o) for o in out], lag.ctypes.data_as(i32p) if want_lag else None,
                                 ctypes.byref(ms))
    if rc != HTM_OK:
        buf = ctypes.create_string_buffer(512)
        lib.htm_last_error(None, buf, 512)
        raise HtmError(rc, buf.value.decode())
    r = dict(t=out[0], t_stdv=out[1], amp=out[2], amp_stdv=out[3], kernel_ms=ms.value)
    if want_lag:
        r["lag"] = lag
    return r


def detect_windows(env, n_smp, n_step, alpha, n_pair_thred, n_win=None, device=0):
    """hypo_tremor_measure's scan_cc with the correlation functions of hypo_tremor_correlate recomputed on the device.
    env: [n_sta, n_total].  Returns dict(cc_thred [n_pair], cc_max [n_pair, n_win], detected [n_win] (bool),
    n_pairs_above [n_win], win_id (1-based ids of the detected windows), kernel_ms)."""
    lib = load_library()
    env = _f64(env)
    S, n_total = env.shape
    if n_win is None:
        n_win = (n_total - n_smp) // n_step                       # src/cls_correlator.f90:80
    P = S * (S - 1) // 2
    thr, mx = np.empty(P), np.empty((P, n_win))
    det, cnt = np.zeros(n_win, dtype=np.int32), np.zeros(n_win, dtype=np.int32)
    ms = ctypes.c_double()
    i32p = ctypes.POINTER(ctypes.c_int32)
    rc = lib.htm_detect_windows(device, S, n_total, _dptr(env), n_smp, n_step, alpha, n_pair_thred, n_win, _dptr(thr), _dptr(mx),
                                det.ctypes.data_as(i32p), cnt.ctypes.data_as(i32p), ctypes.byref(ms))
    if rc != HTM_OK:
        buf = ctypes.create_string_buffer(512)
        lib.htm_last_error(None, buf, 512)
        raise HtmError(rc, buf.value.decode())
    return dict(cc_thred=thr, cc_max=mx, detected=det.astype(bool), n_pairs_above=cnt, win_id=np.nonzero(det)[0] + 1,
                kernel_ms=ms.value)


def measure_fp64_peak(device=0):
    lib = load_library()
    tf = ctypes.c_double()
    rc = lib.htm_measure_fp64_peak(device, ctypes.byref(tf))
    if rc != HTM_OK:
        raise HtmError(rc, "htm_measure_fp64_peak failed (no CUDA device?)")
    return tf.value


class HypoTremorB200:
    """One handle = one CUDA device = one shard of events."""

    def __init__(self, cfg):
        self.lib = load_library()
        self.cfg = copy_config(cfg)
        self._h = ctypes.c_void_p()
        rc = self.lib.htm_create(ctypes.byref(self._h), ctypes.byref(self.cfg))
        if rc != HTM_OK:
            buf = ctypes.create_string_buffer(512)
            self.lib.htm_last_error(None, buf, 512)
            self._h = ctypes.c_void_p()
            raise HtmError(rc, buf.value.decode())
        from .synth import shard_bounds
        self.n_sta, self.n_procs, self.n_chains = cfg.n_sta, cfg.n_procs, cfg.n_chains
        self.rank_offset = 0
        if cfg.mode == MODE_BLOCKED_GIBBS and not cfg.gibbs_shard_events:
            # joint chains: the shards split the virtual ranks and every shard holds all events
            lo, hi = shard_bounds(cfg.n_procs, cfg.shard_rank, cfg.shard_count)
            self.event_offset, self.n_events = 0, cfg.n_events
            self.rank_offset, self.n_procs = lo, hi - lo
        else:
            lo, hi = shard_bounds(cfg.n_events, cfg.shard_rank, cfg.shard_count)
            self.event_offset, self.n_events = lo, hi - lo

    # -- plumbing ------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != HTM_OK:
            buf = ctypes.create_string_buffer(512)
            self.lib.htm_last_error(self._h, buf, 512)
            raise HtmError(rc, buf.value.decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.htm_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- inputs --------------------------------------------------------------------------
    def set_stations(self, sta_x, sta_y, sta_z):
        S = self.n_sta
        x, y, z = _f64(sta_x, (S,)), _f64(sta_y, (S,)), _f64(sta_z, (S,))
        self._ck(self.lib.htm_set_stations(self._h, _dptr(x), _dptr(y), _dptr(z)))

    def set_observations(self, t_obs, t_stdv, a_obs, a_stdv):
        shp = (self.n_events, self.n_sta)
        a = [_f64(v, shp) for v in (t_obs, t_stdv, a_obs, a_stdv)]
        self._ck(self.lib.htm_set_observations(self._h, *[_dptr(v) for v in a]))

    def set_xy_prior(self, x_mu, y_mu):
        x, y = _f64(x_mu, (self.n_events,)), _f64(y_mu, (self.n_events,))
        self._ck(self.lib.htm_set_xy_prior(self._h, _dptr(x), _dptr(y)))

    def set_globals(self, vs, qs, t_corr=None, a_corr=None):
        tc = _f64(t_corr, (self.n_sta,)) if t_corr is not None else None
        ac = _f64(a_corr, (self.n_sta,)) if a_corr is not None else None
        self._ck(self.lib.htm_set_globals(self._h, vs, qs, _dptr(tc), _dptr(ac)))

    def load(self, syn):
        """Convenience: stations, observations and xy prior from a Synthetic (shard)."""
        self.set_stations(syn.sta_x, syn.sta_y, syn.sta_z)
        self.set_observations(syn.t_obs, syn.t_stdv, syn.a_obs, syn.a_stdv)
        self.set_xy_prior(syn.x_mu, syn.y_mu)

    # -- chains --------------------------------------------------------------------------
    def init_chains(self):
        self._ck(self.lib.htm_init_chains(self._h))

    def set_chain_state(self, rank, chain, hypo, t_corr, a_corr, vs, qs, temp, log_likelihood):
        h = _f64(hypo, (3 * self.n_events,))
        tc, ac = _f64(t_corr, (self.n_sta,)), _f64(a_corr, (self.n_sta,))
        self._ck(self.lib.htm_set_chain_state(self._h, rank, chain, _dptr(h), _dptr(tc), _dptr(ac),
                                              vs, qs, temp, log_likelihood))

    def get_chain_state(self, rank, chain):
        h = np.empty(3 * self.n_events)
        tc, ac = np.empty(self.n_sta), np.empty(self.n_sta)
        s = [ctypes.c_double() for _ in range(4)]
        self._ck(self.lib.htm_get_chain_state(self._h, rank, chain, _dptr(h), _dptr(tc), _dptr(ac),
                                              *[ctypes.byref(v) for v in s]))
        return dict(hypo=h, t_corr=tc, a_corr=ac, vs=s[0].value, qs=s[1].value, temp=s[2].value,
                    log_likelihood=s[3].value)

    # -- hot path ------------------------------------------------------------------------
    def loglik(self, hypo, t_corr, a_corr, vs, qs, per_event=False):
        hypo = _f64(hypo)
        M = hypo.shape[0]
        hypo = _f64(hypo, (M, 3 * self.n_events))
        tc, ac = _f64(t_corr, (M, self.n_sta)), _f64(a_corr, (M, self.n_sta))
        vs, qs = _f64(vs, (M,)), _f64(qs, (M,))
        L = np.empty(M)
        pe = np.empty((M, self.n_events)) if per_event else None
        self._ck(self.lib.htm_loglik(self._h, M, _dptr(hypo), _dptr(tc), _dptr(ac), _dptr(vs),
                                     _dptr(qs), _dptr(L), _dptr(pe)))
        return (L, pe) if per_event else L

    def run(self, iter_first, iter_last):
        self._ck(self.lib.htm_run(self._h, iter_first, iter_last))

    def run_traced(self, iter_first, iter_last):
        n_it = iter_last - iter_first + 1
        if self.cfg.mode == MODE_BLOCKED_GIBBS:
            # row n_events of axis 1 is the shared-parameter step; one swap attempt per iteration
            tr = np.zeros((n_it, self.n_events + 1, self.n_procs, self.n_chains), dtype=STEP_TRACE_DTYPE)
            sw = np.zeros(n_it, dtype=SWAP_TRACE_DTYPE)
        else:
            tr = np.zeros((n_it, self.n_events, self.n_procs, self.n_chains), dtype=STEP_TRACE_DTYPE)
            sw = np.zeros((n_it, self.n_events, self.n_procs), dtype=SWAP_TRACE_DTYPE)
        self._ck(self.lib.htm_run_traced(self._h, iter_first, iter_last, tr.ctypes.data, sw.ctypes.data))
        return tr, sw

    def synchronize(self):
        self._ck(self.lib.htm_synchronize(self._h))

    def gibbs_pending(self):
        """(which, idx, x_new) per joint chain: the shared-parameter proposal the next iteration will judge."""
        J = self.n_procs * self.n_chains
        which, idx, xn = np.zeros(J, dtype=np.int32), np.zeros(J, dtype=np.int32), np.zeros(J)
        ip = ctypes.POINTER(ctypes.c_int32)
        self._ck(self.lib.htm_gibbs_pending(self._h, which.ctypes.data_as(ip), idx.ctypes.data_as(ip), _dptr(xn)))
        return which, idx, xn

    def gibbs_last_sums(self):
        """(cur, prop) per joint chain: sums over this shard's events judged by the last iteration (float32 mode C)."""
        J = self.n_procs * self.n_chains
        cur, prop = np.zeros(J), np.zeros(J)
        self._ck(self.lib.htm_gibbs_last_sums(self._h, _dptr(cur), _dptr(prop)))
        return cur, prop

    def replay(self, iter_first, iter_last, draws):
        """draws: list (one per virtual rank) of int32 arrays of raw xorshift128 words."""
        R = self.n_procs
        arrs = [np.ascontiguousarray(d, dtype=np.int32) for d in draws]
        ptrs = (ctypes.POINTER(ctypes.c_int32) * R)(
            *[a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)) for a in arrs])
        n = np.array([a.size for a in arrs], dtype=np.int64)
        used = np.zeros(R, dtype=np.int64)
        n_it = iter_last - iter_first + 1
        tr = np.zeros((n_it, R, self.n_chains), dtype=STEP_TRACE_DTYPE)
        sw = np.zeros(n_it, dtype=SWAP_TRACE_DTYPE)
        self._ck(self.lib.htm_replay(self._h, iter_first, iter_last, ptrs,
                                     n.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                     tr.ctypes.data, sw.ctypes.data,
                                     used.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))))
        return tr, sw, used

    # -- outputs -------------------------------------------------------------------------
    def fetch_samples(self, rank, max_records=1 << 20):
        E, S = self.n_events, self.n_sta
        n = ctypes.c_int32()
        # bounded by the ring capacity (in blocked-Gibbs mode a rank can momentarily hold every cold chain)
        cap = min(max_records, max(1, self.cfg.max_samples) * self.cfg.n_cool * self.n_procs)
        it = np.empty(cap, dtype=np.int32)
        vs, qs = np.empty(cap), np.empty(cap)
        hypo = np.empty((cap, 3 * E))
        tc, ac = np.empty((cap, S)), np.empty((cap, S))
        self._ck(self.lib.htm_fetch_samples(self._h, rank, cap, ctypes.byref(n),
                                            it.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                            _dptr(vs), _dptr(qs), _dptr(hypo), _dptr(tc), _dptr(ac)))
        k = n.value
        return dict(iter=it[:k], vs=vs[:k], qs=qs[:k], hypo=hypo[:k], t_corr=tc[:k], a_corr=ac[:k])

    def fetch_likelihood(self, rank, max_records=1 << 20):
        n = ctypes.c_int32()
        cap = min(max_records, max(1, self.cfg.max_samples) * self.cfg.n_cool * self.cfg.n_procs)
        it = np.empty(cap, dtype=np.int32)
        lik = np.empty(cap)
        self._ck(self.lib.htm_fetch_likelihood(self._h, rank, cap, ctypes.byref(n),
                                               it.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                               _dptr(lik)))
        return it[:n.value], lik[:n.value]

    def posterior_quantiles(self):
        """dict of {median, 2.5 %, 97.5 %} per marginal from the device-side store (cfg.summary = 1):
        hypo [3 E, 3], vs [3], qs [3], t_corr [S, 3], a_corr [S, 3], n."""
        E, S = self.n_events, self.n_sta
        n = ctypes.c_int32()
        hq, vq, qq = np.empty((3 * E, 3)), np.empty(3), np.empty(3)
        tq, aq = np.empty((S, 3)), np.empty((S, 3))
        self._ck(self.lib.htm_posterior_quantiles(self._h, ctypes.byref(n), _dptr(hq), _dptr(vq), _dptr(qq), _dptr(tq),
                                                  _dptr(aq)))
        return dict(n=n.value, hypo=hq, vs=vq, qs=qq, t_corr=tq, a_corr=aq)

    def discard_samples(self):
        self._ck(self.lib.htm_discard_samples(self._h))

    def get_counts(self):
        p, a = np.zeros(7, dtype=np.int64), np.zeros(7, dtype=np.int64)
        lp = ctypes.POINTER(ctypes.c_int64)
        self._ck(self.lib.htm_get_counts(self._h, p.ctypes.data_as(lp), a.ctypes.data_as(lp)))
        return p, a

    def get_histograms(self):
        h = np.zeros((self.n_events, 3, self.cfg.hist_bins), dtype=np.uint32)
        self._ck(self.lib.htm_get_histograms(self._h, h.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))))
        return h

    # -- multi-GPU at the ABI level (NCCL inside the library) ----------------------------------
    @staticmethod
    def comm_unique_id():
        """128-byte NCCL id; create on one shard, hand to every shard."""
        lib = load_library()
        buf = ctypes.create_string_buffer(128)
        rc = lib.htm_comm_unique_id(buf)
        if rc != HTM_OK:
            err = ctypes.create_string_buffer(512)
            lib.htm_last_error(None, err, 512)
            raise HtmError(rc, err.value.decode())
        return buf.raw

    def comm_init(self, unique_id):
        self._ck(self.lib.htm_comm_init(self._h, ctypes.create_string_buffer(unique_id, 128)))

    def gather_samples(self, rank, max_records=1 << 20):
        """Collective fetch_samples (event-sharded factorised mode): hypocentres of ALL events on every shard."""
        E_tot, S = self.cfg.n_events, self.n_sta
        n = ctypes.c_int32()
        cap = min(max_records, max(1, self.cfg.max_samples) * self.cfg.n_cool * self.n_procs)
        it = np.empty(cap, dtype=np.int32)
        vs, qs = np.empty(cap), np.empty(cap)
        hypo = np.empty((cap, 3 * E_tot))
        tc, ac = np.empty((cap, S)), np.empty((cap, S))
        self._ck(self.lib.htm_gather_samples(self._h, rank, cap, ctypes.byref(n),
                                             it.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                             _dptr(vs), _dptr(qs), _dptr(hypo), _dptr(tc), _dptr(ac)))
        k = n.value
        return dict(iter=it[:k], vs=vs[:k], qs=qs[:k], hypo=hypo[:k], t_corr=tc[:k], a_corr=ac[:k])

    def comm_p2p_export(self):
        """64-byte CUDA IPC handle of this shard's exchange buffer (event-sharded blocked Gibbs)."""
        buf = ctypes.create_string_buffer(64)
        self._ck(self.lib.htm_comm_p2p_export(self._h, buf))
        return buf.raw

    def comm_p2p_import(self, handles):
        """handles: the 64-byte handles of all shards in shard order (the host program all-gathers them)."""
        blob = b"".join(handles)
        assert len(blob) == 64 * self.cfg.shard_count
        self._ck(self.lib.htm_comm_p2p_import(self._h, ctypes.create_string_buffer(blob, len(blob))))

    def gather(self, histograms=True):
        """(hist_all [n_events_total, 3, bins] or None, n_propose[7], n_accept[7]) over all shards."""
        hist = np.zeros((self.cfg.n_events, 3, self.cfg.hist_bins), dtype=np.uint32) if histograms else None
        p, a = np.zeros(7, dtype=np.int64), np.zeros(7, dtype=np.int64)
        lp = ctypes.POINTER(ctypes.c_int64)
        hp = hist.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)) if histograms else None
        self._ck(self.lib.htm_gather(self._h, hp, p.ctypes.data_as(lp), a.ctypes.data_as(lp)))
        return hist, p, a

    def device_ptr(self, what):
        p, n = ctypes.c_void_p(), ctypes.c_int64()
        self._ck(self.lib.htm_device_ptr(self._h, what, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def last_run_stats(self):
        ms, nl, npr = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int64()
        self._ck(self.lib.htm_last_run_stats(self._h, ctypes.byref(ms), ctypes.byref(nl), ctypes.byref(npr)))
        return ms.value, nl.value, npr.value
