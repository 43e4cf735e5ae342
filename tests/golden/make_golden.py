"""Generates tests/golden/forward_golden.npz: inputs and log-likelihoods from an INDEPENDENT
numpy restatement of src/cls_forward.f90 (vectorised, written separately from the C++ oracle).

The reference ships no golden vectors and cannot be compiled in this image (no Fortran
compiler), so these vectors pin the oracle against a second, differently-structured statement
of the same formulas rather than against reference output: "parity unpinned" by the reference
itself.  Also stores the hand-derived mod_random known answers of SURVEY.md section 8a.

    python tests/golden/make_golden.py
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LOG_2PI_HALF = 0.5 * np.log(2.0 * np.arccos(-1.0))  # src/cls_forward.f90:5


def forward_tables(t_stdv, a_stdv):
    """init_forward, src/cls_forward.f90:76-92: the branch looks at t_stdv only."""
    ok = t_stdv > 1e-16
    ts = np.where(ok, t_stdv, 1.0)
    as_ = np.where(ok, a_stdv, 1.0)
    log_t = np.where(ok, np.log(np.where(ok, t_stdv, 1.0)), 1.0)
    log_a = np.where(ok, np.log(np.where(ok, a_stdv, 1.0)), 1.0)
    return ts, as_, 1.0 / ts ** 2, 1.0 / as_ ** 2, log_t, log_a


def loglik(sta, hypo, t_corr, a_corr, vs, qs, t_obs, t_stdv, a_obs, a_stdv, use_time=True, use_amp=True):
    """forward_calc_log_likelihood (:268-303) with calc_travel_time (:100-138) and calc_amp
    (:183-222).  Arrays (E, S).  Returns (L, per_event)."""
    ts, as_, wt, wa, log_t, log_a = forward_tables(t_stdv, a_stdv)
    h = hypo.reshape(-1, 3)
    d = np.sqrt((h[:, 0:1] - sta[0][None, :]) ** 2 + (h[:, 1:2] - sta[1][None, :]) ** 2
                + (h[:, 2:3] - sta[2][None, :]) ** 2)
    per_event = np.zeros(h.shape[0])
    if use_time:
        t_syn = d / vs - t_corr[None, :]
        t_mean = np.sum(wt * (t_syn - t_obs), axis=1) / np.sum(wt, axis=1)
        t_syn = t_syn - t_mean[:, None]
        per_event += np.sum(-(t_obs - t_syn) ** 2 / (2.0 * ts ** 2) - LOG_2PI_HALF - log_t, axis=1)
    if use_amp:
        a_syn = -d * np.pi * 5.0 / (qs * vs) - np.log(d) - a_corr[None, :]
        a_mean = np.sum(wa * (a_syn - a_obs), axis=1) / np.sum(wa, axis=1)
        a_syn = a_syn - a_mean[:, None]
        per_event += np.sum(-(a_obs - a_syn) ** 2 / (2.0 * as_ ** 2) - LOG_2PI_HALF - log_a, axis=1)
    return float(np.sum(per_event)), per_event


def main():
    rng = np.random.default_rng(424242)
    E, S, M = 7, 13, 6
    sta = np.stack([rng.uniform(-50, 50, S), rng.uniform(-50, 50, S), rng.uniform(0, 3, S)])
    t_obs = rng.normal(0, 4, (E, S))
    a_obs = rng.normal(0, 1, (E, S))
    t_stdv = rng.uniform(0.2, 0.6, (E, S))
    a_stdv = rng.uniform(0.1, 0.3, (E, S))
    # degenerate sigmas: exercises the t-only branch and log sigma := 1.0 (quirk Q2)
    t_stdv[2, 3] = 0.0
    t_stdv[5, 0] = 1e-17
    a_stdv[2, 3] = 0.123          # ignored: branch is decided by t_stdv
    hypo = np.stack([np.stack([rng.uniform(-30, 30, E), rng.uniform(-30, 30, E), rng.uniform(5, 15, E)],
                              axis=1).ravel() for _ in range(M)])
    t_corr = rng.normal(0, 0.3, (M, S))
    a_corr = rng.normal(0, 0.02, (M, S))
    vs = rng.uniform(2.5, 3.5, M)
    qs = rng.uniform(150, 400, M)
    out = dict(sta=sta, t_obs=t_obs, t_stdv=t_stdv, a_obs=a_obs, a_stdv=a_stdv, hypo=hypo, t_corr=t_corr,
               a_corr=a_corr, vs=vs, qs=qs)
    for tag, (ut, ua) in dict(both=(True, True), time=(True, False), amp=(False, True)).items():
        L = np.zeros(M)
        pe = np.zeros((M, E))
        for m in range(M):
            L[m], pe[m] = loglik(sta, hypo[m], t_corr[m], a_corr[m], vs[m], qs[m], t_obs, t_stdv, a_obs, a_stdv,
                                 ut, ua)
        out["L_" + tag] = L
        out["per_event_" + tag] = pe
    # mod_random known answers, hand-derived in SURVEY.md section 8a from src/mod_random.f90:49-52,63-72
    out["rng_seeds"] = np.array([[1267245926, 454128444, 158352566, 6778530],
                                 [823976407, 1820592774, 673410143, 27175005],
                                 [-1124417450, -178805092, 1709616678, 61439730],
                                 [50098311, -1221904538, -761329265, 109978605]], dtype=np.int64)
    out["rng_rank0_rand_u"] = np.array([0.5585087749641389, 0.12064291047863662, 0.582958621205762,
                                        0.6800179961137474])
    np.savez(os.path.join(HERE, "forward_golden.npz"), **out)
    print("wrote forward_golden.npz")


if __name__ == "__main__":
    main()
