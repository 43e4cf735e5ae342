// Forward model + per-event log-likelihood on the device.
//
// What it computes (reference: src/cls_forward.f90):
//   d_j   = |h - X_j|                                            (:157-159, :244-246)
//   tau_j = d_j/beta - tc_j                                      (:157-160)
//   alp_j = -d_j*pi*f/(Q*beta) - ln d_j - ac_j,  f = 5 Hz         (:247)
//   both are demeaned with the precision-weighted mean of (syn - obs)  (:166-173, :252-259)
//   L_e   = - sum_j (obs-syn)^2/(2 sigma^2) - sum_j (0.5 ln 2pi + ln sigma)   (:283-297)
//
// Device formulation.  With r_j = syn_raw_j - obs_j and w_j = sigma_j^-2 the demeaned
// misfit is  chi2 = sum w (r-m)^2,  m = sum w r / sum w.  It is accumulated in ONE pass
// around a shift c (the residual of the event's first station):
//   S1 = sum w (r-c),  S2 = sum w (r-c)^2,  chi2 = S2 - S1^2 / sum w
// The shift removes the large common offset of r (the unknown origin time / source
// amplitude), so the subtraction loses ~1 digit instead of ~4 in float32.
//
// Tables (built on the host by htm_tables.cpp, with the degenerate-sigma rule of
// src/cls_forward.f90:78-90 applied):
//   sta4[S]      = {X, Y, Z, 0}
//   obs4[E][S]   = {t_obs (+tc_j when globals are fixed), w_t, a_obs (+ac_j), w_a}
//                  w_t = 0 when !use_time, w_a = 0 when !use_amp
//   evc4[E]      = {C_e, 1/sum w_t, 1/sum w_a, 0},  C_e = sum_j used (0.5 ln 2pi + ln sigma)
#pragma once
#include "htm_common.cuh"

namespace htm {

// per-chain scalars of the forward model
template <typename real>
struct Glob {
  real beta;   // vs
  real ivs;    // 1/vs
  real qbeta;  // qs*vs
  real B;      // pi*f/(qs*vs)
};
template <typename real>
__device__ __forceinline__ Glob<real> make_glob(real vs, real qs) {
  Glob<real> g;
  g.beta = vs;
  g.ivs = static_cast<real>(1) / vs;  // once per launch: exact division, not the MUFU approximation
  g.qbeta = qs * vs;
  g.B = static_cast<real>(kPi * kFreq) / g.qbeta;
  return g;
}

// raw residuals (syn - obs) of one station; tc/ac are extra per-chain station terms
// (0 when they are folded into the table)
__device__ __forceinline__ void station_resid(float px, float py, float pz, const Glob<float>& g,
                                              const float4 st, const float4 ob, float tc, float ac,
                                              float& rt, float& ra) {
  const float dx = px - st.x, dy = py - st.y, dz = pz - st.z;
  const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
  float d, lnd;
  M<float>::dist(d2, d, lnd);
  rt = fmaf(d, g.ivs, -tc) - ob.x;
  ra = fmaf(-g.B, d, -lnd) - ac - ob.z;
}
// float64: the reference's own operation order (division by beta, d*pi*f/(q*beta))
__device__ __forceinline__ void station_resid(double px, double py, double pz, const Glob<double>& g,
                                              const double4 st, const double4 ob, double tc, double ac,
                                              double& rt, double& ra) {
  const double dx = px - st.x, dy = py - st.y, dz = pz - st.z;
  const double d = ::sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
  rt = (d / g.beta - tc) - ob.x;
  ra = (-d * kPi * kFreq / g.qbeta - ::log(d) - ac) - ob.z;
}

// One station's contribution to the shifted sums S1t = sum w_t (r_t - c_t), S1a, and
// S2 = sum w_t (r_t - c_t)^2 + w_a (r_a - c_a)^2, given nct = -c_t, nca = -c_a.
// float: the shift rides in the FMA addend (18 FP + 2 MUFU per station).
__device__ __forceinline__ void station_accum(float px, float py, float pz, const Glob<float>& g, float nct,
                                              float nca, const float4 st, const float4 ob, float& S1t, float& S1a,
                                              float& S2) {
  const float dx = px - st.x, dy = py - st.y, dz = pz - st.z;
  const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
  const float d = d2 * mufu_rsq(d2);
  const float l2 = mufu_lg2(d2);
  const float et = fmaf(d, g.ivs, nct) - ob.x;
  const float ea = fmaf(-0.34657359027997264f, l2, fmaf(-g.B, d, nca)) - ob.z;
  const float qt = ob.y * et, qa = ob.w * ea;
  S1t += qt;
  S1a += qa;
  S2 = fmaf(qt, et, S2);
  S2 = fmaf(qa, ea, S2);
}
__device__ __forceinline__ void station_accum(double px, double py, double pz, const Glob<double>& g, double nct,
                                              double nca, const double4 st, const double4 ob, double& S1t,
                                              double& S1a, double& S2) {
  double rt, ra;
  station_resid(px, py, pz, g, st, ob, 0.0, 0.0, rt, ra);
  const double et = rt + nct, ea = ra + nca;
  const double qt = ob.y * et, qa = ob.w * ea;
  S1t += qt;
  S1a += qa;
  S2 += qt * et;
  S2 += qa * ea;
}

// ---- packed float32 (FFMA2 / FADD2 / FMUL2): two chains of one thread per instruction -------
// Shared-memory record per station, 16 floats, every value duplicated so that one 64-bit
// register pair feeds both chains:  {-X,-X,-Y,-Y} {-Z,-Z,-t,-t} {w_t,w_t,-a,-a} {w_a,w_a,0,0}
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ void packed_geometry(const float2 px, const float2 py, const float2 pz, const float4 r0,
                                                const float4 r1, float2& d, float2& l2) {
  const float2 dx = __fadd2_rn(px, f2(r0.x, r0.y));
  const float2 dy = __fadd2_rn(py, f2(r0.z, r0.w));
  const float2 dz = __fadd2_rn(pz, f2(r1.x, r1.y));
  const float2 d2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
  d = __fmul2_rn(d2, f2(mufu_rsq(d2.x), mufu_rsq(d2.y)));
  l2 = f2(mufu_lg2(d2.x), mufu_lg2(d2.y));
}
template <int NSLOT>
__device__ __forceinline__ void forward_packed(const float4* __restrict__ s_pk, const int S, const float (&nx)[NSLOT],
                                               const float (&ny)[NSLOT], const float (&nz)[NSLOT],
                                               const Glob<float>& g, float (&S1t)[NSLOT], float (&S1a)[NSLOT],
                                               float (&S2)[NSLOT]) {
  constexpr int NP = NSLOT / 2;
  const float2 ivs2 = f2(g.ivs, g.ivs), nB2 = f2(-g.B, -g.B);
  const float2 nc2 = f2(-0.34657359027997264f, -0.34657359027997264f);
  float2 px[NP], py[NP], pz[NP], nct[NP], nca[NP], a1t[NP], a1a[NP], a2[NP];
  {
    const float4 r0 = s_pk[0], r1 = s_pk[1], r2 = s_pk[2];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      px[p] = f2(nx[2 * p], nx[2 * p + 1]);
      py[p] = f2(ny[2 * p], ny[2 * p + 1]);
      pz[p] = f2(nz[2 * p], nz[2 * p + 1]);
      float2 d, l2;
      packed_geometry(px[p], py[p], pz[p], r0, r1, d, l2);
      // shift = raw residual of station 0; keep its negative
      const float2 rt = __ffma2_rn(d, ivs2, f2(r1.z, r1.w));
      const float2 ra = __ffma2_rn(nc2, l2, __ffma2_rn(nB2, d, f2(r2.z, r2.w)));
      nct[p] = f2(-rt.x, -rt.y);
      nca[p] = f2(-ra.x, -ra.y);
      a1t[p] = f2(0.f, 0.f);
      a1a[p] = f2(0.f, 0.f);
      a2[p] = f2(0.f, 0.f);
    }
  }
#pragma unroll 2
  for (int j = 1; j < S; ++j) {
    const float4 r0 = s_pk[4 * j], r1 = s_pk[4 * j + 1], r2 = s_pk[4 * j + 2];
    const float2 wa = *reinterpret_cast<const float2*>(s_pk + 4 * j + 3);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      float2 d, l2;
      packed_geometry(px[p], py[p], pz[p], r0, r1, d, l2);
      const float2 et = __fadd2_rn(__ffma2_rn(d, ivs2, nct[p]), f2(r1.z, r1.w));
      const float2 ea = __fadd2_rn(__ffma2_rn(nc2, l2, __ffma2_rn(nB2, d, nca[p])), f2(r2.z, r2.w));
      const float2 qt = __fmul2_rn(f2(r2.x, r2.y), et), qa = __fmul2_rn(wa, ea);
      a1t[p] = __fadd2_rn(a1t[p], qt);
      a1a[p] = __fadd2_rn(a1a[p], qa);
      a2[p] = __ffma2_rn(qt, et, a2[p]);
      a2[p] = __ffma2_rn(qa, ea, a2[p]);
    }
  }
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    S1t[2 * p] = a1t[p].x;
    S1t[2 * p + 1] = a1t[p].y;
    S1a[2 * p] = a1a[p].x;
    S1a[2 * p + 1] = a1a[p].y;
    S2[2 * p] = a2[p].x;
    S2[2 * p + 1] = a2[p].y;
  }
}

template <typename real>
__device__ __forceinline__ real finish_loglik(real S1t, real S2t, real S1a, real S2a,
                                              const typename M<real>::real4 evc) {
  const real chi2 = (S2t - S1t * S1t * evc.y) + (S2a - S1a * S1a * evc.z);
  return static_cast<real>(-0.5) * chi2 - evc.x;
}

// ---- warp-cooperative: stations strided over the 32 lanes, shuffle reductions --------------
// sta4/obs4 may point to global or shared memory.  tc/ac: per-chain station terms or nullptr.
// (ov_which, ov_idx, ov_val): optional override of one station term (2 = t_corr, 4 = a_corr),
// used by the joint modes to evaluate a proposal without writing it.
template <typename real, typename corr_t>
__device__ __forceinline__ real warp_event_loglik(const typename M<real>::real4* __restrict__ sta4,
                                                  const typename M<real>::real4* __restrict__ obs4,
                                                  const typename M<real>::real4 evc, int S, real px, real py,
                                                  real pz, const Glob<real>& g, const corr_t* __restrict__ tc,
                                                  const corr_t* __restrict__ ac, int ov_which = 0,
                                                  int ov_idx = -1, real ov_val = 0) {
  typedef typename M<real>::real4 real4;
  const int lane = threadIdx.x & 31;
  real ct = 0, ca = 0;
  real S1t = 0, S2t = 0, S1a = 0, S2a = 0;
  // first round peeled to obtain the shift from station 0 (lane 0)
  for (int j0 = 0; j0 < S; j0 += 32) {
    const int j = j0 + lane;
    real rt = 0, ra = 0, wt = 0, wa = 0;
    if (j < S) {
      const real4 st = sta4[j];
      const real4 ob = obs4[j];
      real tcj = tc ? static_cast<real>(tc[j]) : static_cast<real>(0);
      real acj = ac ? static_cast<real>(ac[j]) : static_cast<real>(0);
      if (j == ov_idx) {
        if (ov_which == 2) tcj = ov_val;
        if (ov_which == 4) acj = ov_val;
      }
      station_resid(px, py, pz, g, st, ob, tcj, acj, rt, ra);
      wt = ob.y;
      wa = ob.w;
    }
    if (j0 == 0) {
      ct = __shfl_sync(0xffffffffu, rt, 0);
      ca = __shfl_sync(0xffffffffu, ra, 0);
    }
    const real et = rt - ct, ea = ra - ca;
    S1t += wt * et;
    S2t += wt * et * et;
    S1a += wa * ea;
    S2a += wa * ea * ea;
  }
  S1t = warp_sum(S1t);
  S1a = warp_sum(S1a);
  real S2 = warp_sum(S2t + S2a);
  return finish_loglik<real>(S1t, S2, S1a, static_cast<real>(0), evc);
}

// ---- one thread, sequential over stations (tables in shared memory, warp-broadcast reads) ---
template <typename real>
__device__ __forceinline__ real lane_event_loglik(const typename M<real>::real4* sta4,
                                                  const typename M<real>::real4* obs4,
                                                  const typename M<real>::real4 evc, int S, real px, real py,
                                                  real pz, const Glob<real>& g) {
  typedef typename M<real>::real4 real4;
  real ct, ca;
  station_resid(px, py, pz, g, sta4[0], obs4[0], static_cast<real>(0), static_cast<real>(0), ct, ca);
  real S1t = 0, S2 = 0, S1a = 0;
#pragma unroll 4
  for (int j = 1; j < S; ++j) {
    const real4 st = sta4[j];
    const real4 ob = obs4[j];
    real rt, ra;
    station_resid(px, py, pz, g, st, ob, static_cast<real>(0), static_cast<real>(0), rt, ra);
    const real et = rt - ct, ea = ra - ca;
    const real qt = ob.y * et, qa = ob.w * ea;
    S1t += qt;
    S1a += qa;
    S2 += qt * et;
    S2 += qa * ea;
  }
  return finish_loglik<real>(S1t, S2, S1a, static_cast<real>(0), evc);
}

}  // namespace htm
