// Mode C (blocked Gibbs): what one thread does for one (chain, event) in one iteration -- the hypocentre
// proposal judged on the event's own log-likelihood, then the chain's pending shared-parameter proposal
// evaluated for that event -- and the likelihood evaluation it uses (station by station in the reference's
// order).  Float64 validation path; included by htm_gibbs.cu only (float32: htm_gibbs_f32.cu).
#pragma once
#include "htm_gibbs_decide.cuh"

namespace htm {

// one thread, sequential over stations; obs row and station terms in shared memory
template <typename real>
__device__ __forceinline__ real event_loglik_corr(const typename M<real>::real4* s_sta,
                                                  const typename M<real>::real4* s_obs_row,
                                                  const typename M<real>::real4 evc, int S, real px, real py,
                                                  real pz, const Glob<real>& g, const real* s_tc, const real* s_ac,
                                                  int ov_which, int ov_idx, real ov_val) {
  typedef typename M<real>::real4 real4;
  real ct = 0, ca = 0, S1t = 0, S1a = 0, S2 = 0;
#pragma unroll 2
  for (int j = 0; j < S; ++j) {
    const real4 st = s_sta[j];
    const real4 ob = s_obs_row[j];
    real tc = s_tc[j], ac = s_ac[j];
    if (j == ov_idx) {
      if (ov_which == 2) tc = ov_val;
      if (ov_which == 4) ac = ov_val;
    }
    real rt, ra;
    station_resid(px, py, pz, g, st, ob, tc, ac, rt, ra);
    if (j == 0) {
      ct = rt;
      ca = ra;
    }
    const real et = rt - ct, ea = ra - ca;
    const real qt = ob.y * et, qa = ob.w * ea;
    S1t += qt;
    S1a += qa;
    S2 += qt * et;
    S2 += qa * ea;
  }
  return finish_loglik<real>(S1t, S2, S1a, static_cast<real>(0), evc);
}

// ---- per-thread step shared by the per-iteration sweep and the persistent kernel ---------------------
template <typename real>
struct StepIn {
  real T, iT, vs, qs, pval;
  bool cold;
  int which, pidx;
  int S;
  const typename M<real>::real4* obs_row;  // this lane's event row in shared memory
  const typename M<real>::real4* s_sta;
  const real* tc;
  const real* ac;
};

// 1. one hypocentre coordinate proposed and judged on the event's own log-likelihood with the chain's
//    temperature; 2. the chain's pending shared-parameter proposal evaluated for this event (-> Lp).
template <typename real, bool TRACE>
__device__ __forceinline__ void gibbs_thread_step(const GibbsParams<real>& p, const int it, const StepIn<real>& in,
                                                  const int c, const int e, const int ee, const bool ev_ok,
                                                  const typename M<real>::real4 evc, const real mux, const real muy,
                                                  real& x, real& y, real& z, real& Le, real& Lp, int& icmp, bool& acc,
                                                  htm_step_trace* trace) {
  const Glob<real> g = make_glob<real>(in.vs, in.qs);
  const uint32_t gid = (static_cast<uint32_t>(ee) + p.event_offset) * p.J_total + p.chain_offset + static_cast<uint32_t>(c);
  const u32x4 w = philox4x32_10(p.rk, static_cast<uint32_t>(it), gid, PHX_STEP, 0u);
  icmp = static_cast<int>(below(w.v[0], 3u));
  const real gs = M<real>::gauss(w.v[1], w.v[2]);
  const bool isz = icmp == 0;
  const real x_old = isz ? z : (icmp == 1 ? y : x);
  const real mu = isz ? p.prior_z : (icmp == 1 ? muy : mux);
  const real sigma = isz ? p.width_z : p.width_xy;
  const real step = isz ? p.step_z : p.step_xy;
  const real x_new = x_old + gs * step;
  const real dn = x_new - mu, dl = x_old - mu;
  real lpr = -(dn * dn - dl * dl) / (static_cast<real>(2) * sigma * sigma);
  bool ok = true;
  if (isz) {
    if (x_new <= mu)
      ok = false;
    else
      lpr = lpr + M<real>::log(dn) - M<real>::log(dl);
  }
  const real nx = icmp == 2 ? x_new : x, ny = icmp == 1 ? x_new : y, nz = isz ? x_new : z;
  const real Lnew = event_loglik_corr<real>(in.s_sta, in.obs_row, evc, in.S, nx, ny, nz, g, in.tc, in.ac, 0, -1, 0);
  const real ratio = M<real>::div(Lnew - Le, in.T, in.iT) + lpr;
  const real ru = M<real>::u_co(w.v[3]);
  acc = ok && (ru > static_cast<real>(0)) && (M<real>::log(ru) <= ratio);
  if (acc) {
    x = nx;
    y = ny;
    z = nz;
    Le = Lnew;
  }
  if (TRACE) {
    if (ev_ok && trace) {
      htm_step_trace t;
      t.proposal_type = 5 + icmp;
      t.index = 3 * (e + 1) - icmp;
      t.prior_ok = ok ? 1 : 0;
      t.accepted = acc ? 1 : 0;
      t.log_likelihood = static_cast<double>(Le);
      trace[static_cast<size_t>(e) * p.J + c] = t;
    }
  }
  Lp = Le;
  if (in.which != 0) {
    const Glob<real> gp = make_glob<real>(in.which == 1 ? in.pval : in.vs, in.which == 3 ? in.pval : in.qs);
    Lp = event_loglik_corr<real>(in.s_sta, in.obs_row, evc, in.S, x, y, z, gp, in.tc, in.ac, in.which,
                                 (in.which == 2 || in.which == 4) ? in.pidx : -1, in.pval);
  }
}

}  // namespace htm
