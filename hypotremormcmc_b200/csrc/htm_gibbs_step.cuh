// Mode C (blocked Gibbs): what one thread does for one (chain, event) in one iteration -- the hypocentre
// proposal judged on the event's own log-likelihood, then the chain's pending shared-parameter proposal
// evaluated for that event -- and the two likelihood evaluations it uses (float64: station by station in
// the reference's order; float32: FFMA2-packed station pairs).  Included by htm_gibbs.cu only.
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

// ---- float32 packed evaluation (lane = event, warp = chain) --------------------------------------------
// Row of one event in the expanded table: [0] = A of station 0, [1] = {t_obs0, a_obs0, 0, 0}, then 4 float4
// per station pair (htm_forward.cuh: store_station_pair).  cp[m] = {-tc_j0, -tc_j1, -ac_j0, -ac_j1} are the
// chain's station terms of pair m (warp-broadcast), ntc0/nac0 those of station 0.
__device__ __forceinline__ int gibbs_xrow(int S) { return 2 + 4 * (S / 2); }
constexpr int kGibbsPairUnroll = 2;
__device__ __forceinline__ float eval_pairs_f32(const float4* __restrict__ row, const int n_pairs, const float hx,
                                                const float hy, const float hz, const Glob<float>& g,
                                                const float4* __restrict__ cp, const float ntc0, const float nac0,
                                                const float4 evc) {
  const float kC = 0.34657359027997264f;
  const float4 A0 = row[0], h1 = row[1];
  const float h2 = fmaf(hz, hz, fmaf(hy, hy, hx * hx));
  float2 nct, nca;
  {
    const float d2 = fmaf(hx, A0.x, fmaf(hy, A0.y, fmaf(hz, A0.z, A0.w + h2)));
    const float d = d2 * mufu_rsq(d2);
    const float l2 = mufu_lg2(d2);
    const float ct = -(fmaf(d, g.ivs, ntc0) - h1.x);
    const float ca = -(fmaf(-kC, l2, fmaf(-g.B, d, nac0)) - h1.y);
    nct = f2(ct, ct);
    nca = f2(ca, ca);
  }
  const float2 px = f2(hx, hx), py = f2(hy, hy), pz = f2(hz, hz), hh = f2(h2, h2);
  const float2 ivs2 = f2(g.ivs, g.ivs), nB2 = f2(-g.B, -g.B), nc2 = f2(-kC, -kC);
  float2 a1t = f2(0.f, 0.f), a1a = f2(0.f, 0.f), a2 = f2(0.f, 0.f);
  const float4* r = row + 2;
#pragma unroll kGibbsPairUnroll
  for (int m = 0; m < n_pairs; ++m) {
    const float4 r0 = r[4 * m], r1 = r[4 * m + 1], r2 = r[4 * m + 2], r3 = r[4 * m + 3];
    const float4 c4 = cp[m];
    const float2 swt = f2(r2.x, r2.y), swa = f2(r3.x, r3.y);
    const float2 d2 = __ffma2_rn(px, f2(r0.x, r0.y),
                                 __ffma2_rn(py, f2(r0.z, r0.w), __ffma2_rn(pz, f2(r1.x, r1.y), __fadd2_rn(f2(r1.z, r1.w), hh))));
    const float2 d = __fmul2_rn(d2, f2(mufu_rsq(d2.x), mufu_rsq(d2.y)));
    const float2 l2 = f2(mufu_lg2(d2.x), mufu_lg2(d2.y));
    const float2 at = __fadd2_rn(__ffma2_rn(d, ivs2, nct), f2(c4.x, c4.y));
    const float2 ut = __ffma2_rn(swt, at, f2(r2.z, r2.w));
    const float2 aa = __fadd2_rn(__ffma2_rn(nc2, l2, __ffma2_rn(nB2, d, nca)), f2(c4.z, c4.w));
    const float2 ua = __ffma2_rn(swa, aa, f2(r3.z, r3.w));
    a2 = __ffma2_rn(ut, ut, a2);
    a1t = __ffma2_rn(swt, ut, a1t);
    a2 = __ffma2_rn(ua, ua, a2);
    a1a = __ffma2_rn(swa, ua, a1a);
  }
  return finish_loglik<float>(a1t.x + a1t.y, a2.x + a2.y, a1a.x + a1a.y, 0.f, evc);
}


// ---- per-thread step shared by the per-iteration sweep and the persistent kernel ---------------------
template <typename real>
struct StepIn {
  real T, iT, vs, qs, pval;
  bool cold;
  int which, pidx;
  int S, n_pairs;
  const typename M<real>::real4* obs_row;  // this lane's event row in shared memory
  // float32 operands (station terms per pair: current / proposed; station 0: {-tc0, -ac0, -tc0', -ac0'})
  const float4* cp;
  const float4* cpP;
  float4 c0;
  // float64 operands
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
  constexpr bool kF32 = sizeof(real) == 4;
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
  real Lnew;
  if constexpr (kF32) {
    Lnew = eval_pairs_f32(reinterpret_cast<const float4*>(in.obs_row), in.n_pairs, nx - mux, ny - muy, nz, g, in.cp,
                          in.c0.x, in.c0.y, evc);
  } else {
    Lnew = event_loglik_corr<real>(in.s_sta, in.obs_row, evc, in.S, nx, ny, nz, g, in.tc, in.ac, 0, -1, 0);
  }
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
    if constexpr (kF32) {
      Lp = eval_pairs_f32(reinterpret_cast<const float4*>(in.obs_row), in.n_pairs, x - mux, y - muy, z, gp, in.cpP,
                          in.c0.z, in.c0.w, evc);
    } else {
      Lp = event_loglik_corr<real>(in.s_sta, in.obs_row, evc, in.S, x, y, z, gp, in.tc, in.ac, in.which,
                                   (in.which == 2 || in.which == 4) ? in.pidx : -1, in.pval);
    }
  }
}

}  // namespace htm
