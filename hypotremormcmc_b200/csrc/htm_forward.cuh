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

// ---- float32 throughput formulation (lane-per-chain kernels) --------------------------------------
// The staged tables are expanded once per launch, in shared memory, into a record that needs only
// 14 FMA-pipe operations + 2 MUFU per chain-station:
//   A = {cx, cy, cz, c0}     cx = -2(X - x0), cy = -2(Y - y0), cz = -2 Z, c0 = (X-x0)^2 + (Y-y0)^2 + Z^2
//   B = {sw_t, -sw_t*t_obs, sw_a, -sw_a*a_obs}        sw = sqrt(w) = 1/sigma
// with (x0, y0) the event's prior centre.  For a hypocentre h (centred the same way, hh = |h|^2):
//   d^2 = c0 + hh + h.c                       (expansion around the centre: 1 add + 3 FMA instead of 6 ops;
//                                              centring keeps the cancellation harmless)
//   u_t = sw_t (d/vs - c_t - t_obs)           = fma(sw_t, fma(d, 1/vs, -c_t), -sw_t t_obs)
//   u_a = sw_a (-B d - ln d - c_a - a_obs)
//   S2 += u_t^2 + u_a^2 ;  S1t += sw_t u_t ;  S1a += sw_a u_a       (so S1 = sum w e, S2 = sum w e^2)
struct StaRecF {
  float4 A, B;
};
__device__ __forceinline__ StaRecF expand_station(const float4 st, const float4 ob, float x0, float y0) {
  StaRecF r;
  const float X = st.x - x0, Y = st.y - y0, Z = st.z;
  r.A = make_float4(-2.f * X, -2.f * Y, -2.f * Z, fmaf(X, X, fmaf(Y, Y, Z * Z)));
  const float swt = sqrtf(ob.y), swa = sqrtf(ob.w);
  r.B = make_float4(swt, -swt * ob.x, swa, -swa * ob.z);
  return r;
}
// Packed form (FFMA2 / FADD2 / FMUL2, sm_100a): TWO STATIONS of one chain per instruction, so the
// 14 operations per station become 7 issue slots while every lane still owns one chain.  Stations are
// stored as pairs (0,1), (2,3), ...; an odd tail is padded with a zero-weight copy.  Shared-memory record
// per pair, 4 x float4:
//   {cx0,cx1,cy0,cy1} {cz0,cz1,c0_0,c0_1} {sw_t0,sw_t1,-sw_t t_0,-sw_t t_1} {sw_a0,sw_a1,-sw_a a_0,-sw_a a_1}
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ void store_station_pair(float4* dst, const StaRecF a, const StaRecF b) {
  dst[0] = make_float4(a.A.x, b.A.x, a.A.y, b.A.y);
  dst[1] = make_float4(a.A.z, b.A.z, a.A.w, b.A.w);
  dst[2] = make_float4(a.B.x, b.B.x, a.B.y, b.B.y);
  dst[3] = make_float4(a.B.z, b.B.z, a.B.w, b.B.w);
}
#ifndef HTM_PK_UNROLL
#define HTM_PK_UNROLL 2
#endif
constexpr int kPackedUnroll = HTM_PK_UNROLL;
// hx/hy/hz: hypocentres centred on the event's prior centre; nct/nca: the NEGATED shifts c_t, c_a of each
// chain (the lane kernel carries the weighted mean residual of the chain's accepted state, so no station
// needs special treatment: chi2 = S2 - S1^2/W holds for any shift and S1 stays small).
// Returns S1t = sum w_t (r_t - c_t), S1a, S2 = sum w (r - c)^2 per slot.
// the per-thread operands of the packed station loop: NSLOT hypocentres and their running sums
template <int NSLOT>
struct PairAcc {
  float2 px[NSLOT], py[NSLOT], pz[NSLOT], hh[NSLOT], ct2[NSLOT], ca2[NSLOT], a1t[NSLOT], a1a[NSLOT], a2[NSLOT];
  float2 ivs2, nB2, nc2;
  __device__ __forceinline__ void init(const float (&hx)[NSLOT], const float (&hy)[NSLOT], const float (&hz)[NSLOT],
                                       const Glob<float>& g, const float (&nct)[NSLOT], const float (&nca)[NSLOT]) {
    ivs2 = f2(g.ivs, g.ivs);
    nB2 = f2(-g.B, -g.B);
    nc2 = f2(-0.34657359027997264f, -0.34657359027997264f);
#pragma unroll
    for (int q = 0; q < NSLOT; ++q) {
      const float h2 = fmaf(hz[q], hz[q], fmaf(hy[q], hy[q], hx[q] * hx[q]));
      px[q] = f2(hx[q], hx[q]);
      py[q] = f2(hy[q], hy[q]);
      pz[q] = f2(hz[q], hz[q]);
      hh[q] = f2(h2, h2);
      ct2[q] = f2(nct[q], nct[q]);
      ca2[q] = f2(nca[q], nca[q]);
      a1t[q] = f2(0.f, 0.f);
      a1a[q] = f2(0.f, 0.f);
      a2[q] = f2(0.f, 0.f);
    }
  }
  // one station pair (4 float4 of the packed record) for every slot
  __device__ __forceinline__ void pair(const float4* __restrict__ rec) {
    const float4 r0 = rec[0], r1 = rec[1], r2 = rec[2], r3 = rec[3];
    const float2 swt = f2(r2.x, r2.y), swa = f2(r3.x, r3.y);
#pragma unroll
    for (int q = 0; q < NSLOT; ++q) {
      const float2 d2 = __ffma2_rn(px[q], f2(r0.x, r0.y),
                                   __ffma2_rn(py[q], f2(r0.z, r0.w),
                                              __ffma2_rn(pz[q], f2(r1.x, r1.y), __fadd2_rn(f2(r1.z, r1.w), hh[q]))));
#ifdef HTM_ABL_NOMUFU  // tools/micro/pair_loop.cu only: the loop with its MUFU operations replaced by LOP3
      const float2 d = __fmul2_rn(d2, f2(__int_as_float(__float_as_int(d2.x) ^ 0x100), __int_as_float(__float_as_int(d2.y) ^ 0x100)));
      const float2 l2 = f2(__int_as_float(__float_as_int(d2.x) ^ 0x200), __int_as_float(__float_as_int(d2.y) ^ 0x200));
#else
      // (d = MUFU.SQRT(d2) saves the multiply and was measured: +2.8 % with one chain per lane, -2.4 % with two --
      //  and the two layouts must agree bit for bit, so it is not used; profiles/r2bf_lane_phases.txt)
      const float2 d = __fmul2_rn(d2, f2(mufu_rsq(d2.x), mufu_rsq(d2.y)));
      const float2 l2 = f2(mufu_lg2(d2.x), mufu_lg2(d2.y));
#endif
      const float2 ut = __ffma2_rn(swt, __ffma2_rn(d, ivs2, ct2[q]), f2(r2.z, r2.w));
      const float2 ua = __ffma2_rn(swa, __ffma2_rn(nc2, l2, __ffma2_rn(nB2, d, ca2[q])), f2(r3.z, r3.w));
      a2[q] = __ffma2_rn(ut, ut, a2[q]);
      a1t[q] = __ffma2_rn(swt, ut, a1t[q]);
      a2[q] = __ffma2_rn(ua, ua, a2[q]);
      a1a[q] = __ffma2_rn(swa, ua, a1a[q]);
    }
  }
  __device__ __forceinline__ void finish(float (&S1t)[NSLOT], float (&S1a)[NSLOT], float (&S2)[NSLOT]) const {
#pragma unroll
    for (int q = 0; q < NSLOT; ++q) {
      S1t[q] = a1t[q].x + a1t[q].y;
      S1a[q] = a1a[q].x + a1a[q].y;
      S2[q] = a2[q].x + a2[q].y;
    }
  }
};
template <int NSLOT>
__device__ __forceinline__ void forward_pairs(const float4* __restrict__ s_pk, const int n_pairs,
                                              const float (&hx)[NSLOT], const float (&hy)[NSLOT],
                                              const float (&hz)[NSLOT], const Glob<float>& g,
                                              const float (&nct)[NSLOT], const float (&nca)[NSLOT],
                                              float (&S1t)[NSLOT], float (&S1a)[NSLOT], float (&S2)[NSLOT]) {
  PairAcc<NSLOT> A;
  A.init(hx, hy, hz, g, nct, nca);
#pragma unroll kPackedUnroll
  for (int m = 0; m < n_pairs; ++m) A.pair(s_pk + 4 * m);
  A.finish(S1t, S1a, S2);
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
