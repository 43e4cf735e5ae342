"""TEST INFRASTRUCTURE -- ctypes wrapper of oracle/libhtm_oracle.so (the C++ restatement of
the reference hot path).  Never imported by the product package."""
import ctypes
import os
import subprocess

import numpy as np

from hypotremormcmc_b200.config import (HtmConfig, copy_config, STEP_TRACE_DTYPE, SWAP_TRACE_DTYPE,
                                        MODE_FACTORISED, MODE_BLOCKED_GIBBS)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "libhtm_oracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libhtm_oracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.hto_create.restype = ctypes.c_void_p
        L.hto_partial_update.restype = ctypes.c_double
        L.hto_perturb.restype = ctypes.c_double
        L.hto_run_threaded.restype = ctypes.c_double
        L.hto_n_draws.restype = ctypes.c_int64
        _LIB = L
    return _LIB


def _d(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if a is not None else None


def rng_seeds(rank):
    out = (ctypes.c_int32 * 4)()
    lib().hto_rng_seeds(rank, out)
    return list(out)


def rng_draw(rank, kind, n):
    """kind: 0 rand_u, 1 rand_u2, 2 rand_g, 3 rand_r, 4 raw w"""
    out = np.empty(n)
    lib().hto_rng_draw(rank, kind, n, _d(out))
    return out


def philox(seed, c0, c1, c2, c3):
    out = (ctypes.c_uint32 * 4)()
    lib().hto_philox(ctypes.c_uint64(seed), ctypes.c_uint32(c0), ctypes.c_uint32(c1), ctypes.c_uint32(c2),
                     ctypes.c_uint32(c3), out)
    return list(out)


def perturb(x_old, mu, sigma, step, prior_type, g):
    lpr, ok = ctypes.c_double(), ctypes.c_int32()
    f = lib().hto_perturb
    f.argtypes = [ctypes.c_double] * 4 + [ctypes.c_int32, ctypes.c_double, ctypes.POINTER(ctypes.c_double),
                                          ctypes.POINTER(ctypes.c_int32)]
    xn = f(x_old, mu, sigma, step, prior_type, g, ctypes.byref(lpr), ctypes.byref(ok))
    return xn, lpr.value, bool(ok.value)


def judge_swap(t1, t2, l1, l2, r):
    f = lib().hto_judge_swap
    f.argtypes = [ctypes.c_double] * 5
    return bool(f(t1, t2, l1, l2, r))


def select_events(sta_x, sta_y, sta_z, z_guess, t, t_err, a, a_err, vs_min=2.0, vs_max=4.0, b_min=0.015, b_max=0.03):
    """hypo_tremor_select restated (oracle/htm_oracle_select.hpp): dict(vs, t0, b, a0, cc_t, cc_a, selected)"""
    t = np.ascontiguousarray(t, dtype=np.float64)
    E, S = t.shape
    arr = [np.ascontiguousarray(v, dtype=np.float64) for v in (sta_x, sta_y, sta_z, t, t_err, a, a_err)]
    out = np.empty((E, 6))
    sel = np.zeros(E, dtype=np.int32)
    f = lib().hto_select
    dp = ctypes.POINTER(ctypes.c_double)
    f.argtypes = [ctypes.c_int32, ctypes.c_int32, dp, dp, dp, ctypes.c_double, dp, dp, dp, dp] + [ctypes.c_double] * 4 + \
                 [dp, ctypes.POINTER(ctypes.c_int32)]
    f.restype = None
    f(S, E, _d(arr[0]), _d(arr[1]), _d(arr[2]), z_guess, _d(arr[3]), _d(arr[4]), _d(arr[5]), _d(arr[6]), vs_min, vs_max,
      b_min, b_max, _d(out), sel.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    return dict(vs=out[:, 0], t0=out[:, 1], b=out[:, 2], a0=out[:, 3], cc_t=out[:, 4], cc_a=out[:, 5], selected=sel)


def measure_windows(env, dt, n_smp, n_step, win_id, want_lag=False):
    """hypo_tremor_measure's lag / amplitude optimisation restated (oracle/htm_oracle_measure.hpp):
    dict(t, t_stdv, amp, amp_stdv [n_win][S] and, if asked for, lag [n_win][S (S - 1) / 2])"""
    env = np.ascontiguousarray(env, dtype=np.float64)
    S, n_total = env.shape
    win_id = np.ascontiguousarray(win_id, dtype=np.int32)
    W = win_id.size
    out = [np.empty((W, S)) for _ in range(4)]
    lag = np.zeros((W, S * (S - 1) // 2), dtype=np.int32) if want_lag else None
    f = lib().hto_measure
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)
    f.argtypes = [ctypes.c_int32, ctypes.c_int64, dp, ctypes.c_double, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ip,
                  dp, dp, dp, dp, ip]
    f.restype = None
    f(S, n_total, _d(env), dt, n_smp, n_step, W, win_id.ctypes.data_as(ip), *[_d(o) for o in out],
      lag.ctypes.data_as(ip) if want_lag else None)
    r = dict(t=out[0], t_stdv=out[1], amp=out[2], amp_stdv=out[3])
    if want_lag:
        r["lag"] = lag
    return r


def detect_windows(env, n_smp, n_step, alpha, n_pair_thred, n_win=None):
    """scan_cc on recomputed correlation functions (oracle/htm_oracle_measure.hpp: detect_windows)"""
    env = np.ascontiguousarray(env, dtype=np.float64)
    S, n_total = env.shape
    if n_win is None:
        n_win = (n_total - n_smp) // n_step
    P = S * (S - 1) // 2
    thr, mx = np.empty(P), np.empty((P, n_win))
    det, cnt = np.zeros(n_win, dtype=np.int32), np.zeros(n_win, dtype=np.int32)
    f = lib().hto_detect
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)
    f.argtypes = [ctypes.c_int32, ctypes.c_int64, dp, ctypes.c_int32, ctypes.c_int32, ctypes.c_double, ctypes.c_int32,
                  ctypes.c_int32, dp, dp, ip, ip]
    f.restype = None
    f(S, n_total, _d(env), n_smp, n_step, alpha, n_pair_thred, n_win, _d(thr), _d(mx), det.ctypes.data_as(ip), cnt.ctypes.data_as(ip))
    return dict(cc_thred=thr, cc_max=mx, detected=det.astype(bool), n_pairs_above=cnt, win_id=np.nonzero(det)[0] + 1)


class Oracle:
    def __init__(self, cfg, syn, event_offset=0):
        self.L = lib()
        self.cfg = copy_config(cfg)
        self.cfg.n_events = syn.n_events
        self.E, self.S = syn.n_events, syn.n_sta
        self.R, self.K = cfg.n_procs, cfg.n_chains
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in
                (syn.sta_x, syn.sta_y, syn.sta_z, syn.t_obs, syn.t_stdv, syn.a_obs, syn.a_stdv,
                 syn.x_mu, syn.y_mu)]
        self._keep = arrs
        self.h = ctypes.c_void_p(self.L.hto_create(ctypes.byref(self.cfg), *[_d(a) for a in arrs]))
        if event_offset:
            self.L.hto_set_event_offset(self.h, ctypes.c_int32(event_offset))

    def close(self):
        if self.h:
            self.L.hto_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_rank_shard(self, chain_offset, j_total, swap_stream):
        """blocked-Gibbs mode: this oracle holds the virtual ranks of one shard (global Philox ids)"""
        self.L.hto_set_rank_shard(self.h, ctypes.c_uint32(chain_offset), ctypes.c_uint32(j_total),
                                  ctypes.c_uint32(swap_stream))

    def set_sum_hook(self, fn):
        """blocked-Gibbs mode, event shards of one joint ensemble: fn(cur, prop) -> (cur, prop) summed over all
        shards; called once per chain and iteration, in chain order (so a collective inside fn lines up)"""
        proto = ctypes.CFUNCTYPE(None, ctypes.POINTER(ctypes.c_double), ctypes.c_void_p)

        def tramp(two, _user):
            a, b = fn(two[0], two[1])
            two[0], two[1] = a, b

        self._sum_hook = proto(tramp)  # keep the trampoline alive
        self.L.hto_set_sum_hook.argtypes = [ctypes.c_void_p, proto, ctypes.c_void_p]
        self.L.hto_set_sum_hook(self.h, self._sum_hook, None)

    def set_globals(self, vs, qs, t_corr, a_corr):
        tc = np.ascontiguousarray(t_corr, dtype=np.float64)
        ac = np.ascontiguousarray(a_corr, dtype=np.float64)
        self.L.hto_set_globals(self.h, ctypes.c_double(vs), ctypes.c_double(qs), _d(tc), _d(ac))

    def loglik(self, hypo, t_corr, a_corr, vs, qs, per_event=False):
        hypo = np.ascontiguousarray(hypo, dtype=np.float64)
        M = hypo.shape[0]
        tc = np.ascontiguousarray(t_corr, dtype=np.float64)
        ac = np.ascontiguousarray(a_corr, dtype=np.float64)
        vs = np.ascontiguousarray(vs, dtype=np.float64)
        qs = np.ascontiguousarray(qs, dtype=np.float64)
        L = np.empty(M)
        pe = np.empty((M, self.E)) if per_event else None
        self.L.hto_loglik(self.h, ctypes.c_int32(M), _d(hypo), _d(tc), _d(ac), _d(vs), _d(qs), _d(L), _d(pe))
        return (L, pe) if per_event else L

    def partial_update(self, evt_id, hypo_old, ll_old, hypo_new, t_corr, vs, a_corr, qs):
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (hypo_old, hypo_new, t_corr, a_corr)]
        return self.L.hto_partial_update(self.h, ctypes.c_int32(evt_id), _d(a[0]), ctypes.c_double(ll_old),
                                         _d(a[1]), _d(a[2]), ctypes.c_double(vs), _d(a[3]), ctypes.c_double(qs))

    def forward_tables(self):
        n = (self.E, self.S)
        t = [np.empty(n) for _ in range(4)]
        self.L.hto_forward_tables(self.h, *[_d(a) for a in t])
        return dict(t_precision=t[0], log_t_stdv=t[1], a_precision=t[2], log_a_stdv=t[3])

    def record_draws(self, on=True):
        self.L.hto_record_draws(self.h, ctypes.c_int32(1 if on else 0))

    def init_chains(self):
        self.L.hto_init_chains(self.h)

    def get_chain_state(self, rank, chain):
        h = np.empty(3 * self.E)
        tc, ac = np.empty(self.S), np.empty(self.S)
        s = [ctypes.c_double() for _ in range(4)]
        self.L.hto_get_chain_state(self.h, ctypes.c_int32(rank), ctypes.c_int32(chain), _d(h), _d(tc), _d(ac),
                                   *[ctypes.byref(v) for v in s])
        return dict(hypo=h, t_corr=tc, a_corr=ac, vs=s[0].value, qs=s[1].value, temp=s[2].value,
                    log_likelihood=s[3].value)

    def factorised_state(self):
        shp = (self.E, self.R, self.K)
        a = [np.empty(shp) for _ in range(5)]
        self.L.hto_get_factorised_state(self.h, *[_d(v) for v in a])
        return dict(x=a[0], y=a[1], z=a[2], L=a[3], T=a[4])

    def run(self, iter_first, iter_last, trace=True):
        n_it = iter_last - iter_first + 1
        if self.cfg.mode == MODE_FACTORISED:
            tr = np.zeros((n_it, self.E, self.R, self.K), dtype=STEP_TRACE_DTYPE)
            sw = np.zeros((n_it, self.E, self.R), dtype=SWAP_TRACE_DTYPE)
        elif self.cfg.mode == MODE_BLOCKED_GIBBS:
            # row E of axis 1 is the shared-parameter step of each chain
            tr = np.zeros((n_it, self.E + 1, self.R, self.K), dtype=STEP_TRACE_DTYPE)
            sw = np.zeros(n_it, dtype=SWAP_TRACE_DTYPE)
        else:
            tr = np.zeros((n_it, self.R, self.K), dtype=STEP_TRACE_DTYPE)
            sw = np.zeros(n_it, dtype=SWAP_TRACE_DTYPE)
        if trace:
            self.L.hto_run(self.h, ctypes.c_int32(iter_first), ctypes.c_int32(iter_last),
                           ctypes.c_void_p(tr.ctypes.data), ctypes.c_void_p(sw.ctypes.data))
            return tr, sw
        self.L.hto_run(self.h, ctypes.c_int32(iter_first), ctypes.c_int32(iter_last), None, None)
        return None, None

    def run_threaded(self, iter_first, iter_last):
        """mode A with one thread per virtual rank; returns wall seconds"""
        return self.L.hto_run_threaded(self.h, ctypes.c_int32(iter_first), ctypes.c_int32(iter_last))

    def draws(self, rank):
        n = self.L.hto_n_draws(self.h, ctypes.c_int32(rank))
        buf = np.empty(n, dtype=np.int32)
        if n:
            self.L.hto_get_draws(self.h, ctypes.c_int32(rank), buf.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
        return buf

    def clear_draws(self):
        self.L.hto_clear_draws(self.h)

    def get_counts(self):
        p, a = np.zeros(7, dtype=np.int64), np.zeros(7, dtype=np.int64)
        lp = ctypes.POINTER(ctypes.c_int64)
        self.L.hto_get_counts(self.h, p.ctypes.data_as(lp), a.ctypes.data_as(lp))
        return p, a

    def fetch_samples(self, rank):
        n = self.L.hto_n_samples(self.h, ctypes.c_int32(rank))
        cap = max(n, 1)
        it = np.empty(cap, dtype=np.int32)
        vs, qs = np.empty(cap), np.empty(cap)
        hypo = np.empty((cap, 3 * self.E))
        tc, ac = np.empty((cap, self.S)), np.empty((cap, self.S))
        k = ctypes.c_int32()
        self.L.hto_fetch_samples(self.h, ctypes.c_int32(rank), ctypes.c_int32(cap), ctypes.byref(k),
                                 it.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _d(vs), _d(qs), _d(hypo),
                                 _d(tc), _d(ac))
        k = k.value
        return dict(iter=it[:k], vs=vs[:k], qs=qs[:k], hypo=hypo[:k], t_corr=tc[:k], a_corr=ac[:k])

    def fetch_likelihood(self, rank):
        n = self.L.hto_n_likelihood(self.h, ctypes.c_int32(rank))
        cap = max(n, 1)
        it = np.empty(cap, dtype=np.int32)
        lik = np.empty(cap)
        k = ctypes.c_int32()
        self.L.hto_fetch_likelihood(self.h, ctypes.c_int32(rank), ctypes.c_int32(cap), ctypes.byref(k),
                                    it.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _d(lik))
        return it[:k.value], lik[:k.value]
