// Mode B (factorised) kernels: every (event, virtual rank, chain) is an independent tempered
// Metropolis chain over one event's (x, y, z); the K chains of one (event, rank) form a
// tempering group with one swap attempt per iteration.  The whole loop body of
// src/hypo_tremor_mcmc.f90:236-284 (propose -> forward -> judge -> record -> swap) runs
// in-kernel for iter_first..iter_last; chain state stays in registers for the launch.
//
// Per-step rules restated from the reference (the test oracle under oracle/ holds
// the CPU statement of exactly this schedule and is never linked here):
//   component  icmp = int(u*3): 0 -> z, 1 -> y, 2 -> x        src/cls_mcmc.f90:161-163
//   perturb    x' = x + N(0,1)*step; prior ratio               src/cls_model.f90:170-187
//   judge      ln r <= (L'-L)/T + ln prior ratio                src/cls_mcmc.f90:193-203
//   counters   only chains at T = 1                             src/cls_mcmc.f90:186-189,215-218
//   record     T = 1 and mod(it, n_interval) == 1               src/hypo_tremor_mcmc.f90:270-280
//   swap       ln r <= (L2-L1)(1/T1-1/T2); temperatures move    src/cls_parallel.f90:285-302
//
// Two layouts:
//   fact_lane_kernel  one LANE per chain: a warp holds up to 32*NSLOT chains of ONE event, so
//                     every table read is a warp-broadcast from shared memory, no lane idles
//                     whatever n_sta is, sums are thread-local and the swap is a shuffle.
//   fact_warp_kernel  one WARP per chain (the layout BASELINE.json's north_star sketches):
//                     stations across lanes in registers, shuffle-tree reductions, swap
//                     through shared memory + a named barrier per tempering group.
#include <cstdio>
#include <cstdlib>

#include "htm_forward.cuh"
#include "htm_kernels.hpp"

namespace htm {

template <typename real>
struct R2;
template <>
struct R2<float> {
  typedef float2 type;
};
template <>
struct R2<double> {
  typedef double2 type;
};

template <typename real>
struct FactParams {
  typedef typename M<real>::real4 real4;
  typedef typename R2<real>::type real2;
  const real4* sta4;
  const real4* obs4;
  const real4* evc4;
  const real2* prior_xy;
  real *x, *y, *z, *L, *T;
  int E, S, R, K, n_cool;
  int iter_first, iter_last, n_burn, n_interval;
  uint64_t seed;
  PhiloxKeys rk;  // Philox round keys of `seed`
  uint32_t event_offset;
  real vs, qs, prior_z, width_z, width_xy, step_xy, step_z;
  real inv2s2_xy, inv2s2_z;  // 1/(2 sigma^2) of the x,y and z priors (host-computed: no division in the loop)
  // 1/vs and pi f/(qs vs) as make_glob forms them, computed on the host in `real` arithmetic (same bits): values
  // read from the parameter bank are warp-uniform to the compiler, so the packed FFMA2 that use them take a
  // uniform-register operand instead of a third 64-bit register read (tools/micro/issue_mix.cu: 2.1 vs 3.6 cycles)
  real g_ivs, g_B;
  unsigned long long* counts;
  real4* samples;
  int rec_origin, rec_cap;
  uint32_t* hist;
  int hist_bins;
  real hist_hw, hist_zmax;
  htm_step_trace* trace;
  htm_swap_trace* swaps;
};

template <typename real>
__device__ __forceinline__ bool is_cold(real T);
template <>
__device__ __forceinline__ bool is_cold<float>(float T) {
  return T <= 1.0f;  // cold temperatures are exactly 1 and are only ever exchanged
}
template <>
__device__ __forceinline__ bool is_cold<double>(double T) {
  return T < 1.0 + kEps64;  // src/cls_mcmc.f90:186
}

// ---- model_perturb for one hypocentre component (src/cls_model.f90:162-190) -----------------
template <typename real>
__device__ __forceinline__ void propose_hypo(const u32x4& w, real x, real y, real z, real mux, real muy,
                                             const FactParams<real>& p, int& icmp, real& nx, real& ny, real& nz,
                                             real& lpr, bool& ok) {
  icmp = static_cast<int>(below(w.v[0], 3u));
  const real g = M<real>::gauss(w.v[1], w.v[2]);
  const bool isz = icmp == 0;
  const real x_old = isz ? z : (icmp == 1 ? y : x);
  const real mu = isz ? p.prior_z : (icmp == 1 ? muy : mux);
  const real sigma = isz ? p.width_z : p.width_xy;
  const real step = isz ? p.step_z : p.step_xy;
  const real x_new = x_old + g * step;
  const real dn = x_new - mu, dl = x_old - mu;
  lpr = -(dn * dn - dl * dl) / (static_cast<real>(2) * sigma * sigma);
  ok = true;
  if (isz) {
    if (x_new <= mu) {
      ok = false;
    } else {
      lpr = lpr + M<real>::log(dn) - M<real>::log(dl);
    }
  }
  nx = icmp == 2 ? x_new : x;
  ny = icmp == 1 ? x_new : y;
  nz = isz ? x_new : z;
}

// mcmc_judge_model's test (src/cls_mcmc.f90:193-203); u in [0,1) with 24 bits
template <typename real>
__device__ __forceinline__ bool judge(real Lnew, real L, real T, real lpr, bool ok, uint32_t w) {
  const real ratio = (Lnew - L) / T + lpr;
  const real r = M<real>::u_co(w);
  return ok && (r > static_cast<real>(0)) && (M<real>::log(r) <= ratio);
}
// judge_swap (src/cls_parallel.f90:285-302)
template <typename real>
__device__ __forceinline__ bool judge_swap(real T1, real T2, real L1, real L2, uint32_t w) {
  const real del_s = (L2 - L1) * (static_cast<real>(1) / T1 - static_cast<real>(1) / T2);
  const real r = M<real>::u_co(w);
  return (r > static_cast<real>(0)) && (M<real>::log(r) <= del_s);
}

template <typename real>
__device__ __forceinline__ void hist_add(const FactParams<real>& p, int e, real x, real y, real z, real mux,
                                         real muy) {
  const int nb = p.hist_bins;
  const real sxy = static_cast<real>(nb) / (static_cast<real>(2) * p.hist_hw);
  const real sz = static_cast<real>(nb) / p.hist_zmax;
  int bx = static_cast<int>(floor((x - mux + p.hist_hw) * sxy));
  int by = static_cast<int>(floor((y - muy + p.hist_hw) * sxy));
  int bz = static_cast<int>(floor((z - p.prior_z) * sz));
  bx = min(max(bx, 0), nb - 1);
  by = min(max(by, 0), nb - 1);
  bz = min(max(bz, 0), nb - 1);
  uint32_t* h = p.hist + static_cast<size_t>(e) * 3 * nb;
  atomicAdd(h + bx, 1u);
  atomicAdd(h + nb + by, 1u);
  atomicAdd(h + 2 * nb + bz, 1u);
}

// ================================================================================================
// Lane-per-chain kernel
// ================================================================================================
// Per-chain registers: x, y, z, L, T, 1/T, ln(z - prior_z).  Every warp is independent (own
// shared-memory slice, own mbarrier, no block barrier), so the CTA size (1, 2 or 4 warps) is
// chosen by the launcher only to balance warps over the 148 SMs.
#ifndef HTM_LANE_MINB
#define HTM_LANE_MINB 1
#endif
#ifndef HTM_NP_UNROLL
#define HTM_NP_UNROLL 8
#endif
constexpr int kScalarUnroll = HTM_NP_UNROLL;
template <typename real, int NSLOT, bool TRACE>
__global__ void __launch_bounds__(128, HTM_LANE_MINB) fact_lane_kernel(const FactParams<real> p) {
  typedef typename M<real>::real4 real4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int K = p.K, S = p.S, R = p.R;
  const int gpw = min(R, 32 / K);   // tempering groups side by side in one slot row
  const int lanes_used = gpw * K;
  const int gps = gpw * NSLOT;      // groups per warp
  const int wpe = (R + gps - 1) / gps;
  const long gw = static_cast<long>(blockIdx.x) * wpb + warp;
  if (gw >= static_cast<long>(p.E) * wpe) return;  // whole warp leaves; no block barrier below
  const int e = static_cast<int>(gw / wpe), we = static_cast<int>(gw % wpe);

  // --- stage this event's tables with 1-D bulk TMA into the warp's slice of shared memory ---
  // float32: 2 staged + 2 expanded float4 per station (pairs of stations, 4 float4 per pair);
  // float64 reads the staged tables directly
  constexpr bool kF32 = sizeof(real) == 4;
  constexpr int kF4PerSta = kF32 ? 4 : 2;  // (+2 float4 per warp for an odd station count, see launcher)
  const size_t warp_f4 = static_cast<size_t>(kF4PerSta) * S + (kF32 ? 2 : 0);
  real4* s_sta = reinterpret_cast<real4*>(smem_raw) + static_cast<size_t>(warp) * warp_f4;
  real4* s_obs = s_sta + S;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(wpb) * warp_f4 * sizeof(real4)) + warp;
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
    fence_proxy_async();
    const uint32_t bytes = static_cast<uint32_t>(S * sizeof(real4));
    mbar_expect_tx(bar, 2 * bytes);
    tma_load_1d(s_sta, p.sta4, bytes, bar);
    tma_load_1d(s_obs, p.obs4 + static_cast<size_t>(e) * S, bytes, bar);
  }
  __syncwarp();

  // --- chain geometry and state (overlaps the TMA) ---
  const bool lane_ok = lane < lanes_used;
  const int gl = lane_ok ? lane / K : 0, k = lane_ok ? lane % K : 0;
  const int base = gl * K;
  const uint32_t gmask = (K == 32 ? 0xffffffffu : ((1u << K) - 1u)) << base;
  const uint32_t below_me = (1u << lane) - 1u;
  const real4 evc = p.evc4[e];
  const typename R2<real>::type pxy = p.prior_xy[e];
  Glob<real> g;
  g.beta = p.vs;
  g.ivs = p.g_ivs;
  g.qbeta = p.qs * p.vs;
  g.B = p.g_B;
  const uint32_t eg = static_cast<uint32_t>(e) + p.event_offset;
  bool valid[NSLOT];
  int rr[NSLOT];
  size_t ci[NSLOT];
  real x[NSLOT], y[NSLOT], z[NSLOT], L[NSLOT], T[NSLOT], iT[NSLOT], lgz[NSLOT];
  uint32_t cnt_p[3] = {0, 0, 0}, cnt_a[3] = {0, 0, 0};
  uint32_t pk_p = 0, pk_a = 0;  // packed per-type counters (three 10-bit fields) of the last pk_it iterations
  int pk_it = 0;
#pragma unroll
  for (int q = 0; q < NSLOT; ++q) {
    const int r = (we * NSLOT + q) * gpw + gl;
    valid[q] = lane_ok && r < R;
    rr[q] = min(r, R - 1);
    ci[q] = (static_cast<size_t>(e) * R + rr[q]) * K + k;
    x[q] = p.x[ci[q]];
    y[q] = p.y[ci[q]];
    z[q] = p.z[ci[q]];
    L[q] = p.L[ci[q]];
    T[q] = p.T[ci[q]];
    iT[q] = static_cast<real>(1) / T[q];
    lgz[q] = M<real>::log(z[q] - p.prior_z);
  }
  // swap draws of a group: lane k of the group owns iteration blk*K + k + 1 (Philox is counter
  // based, so the words are the same whichever lane evaluates them); refilled every K iterations
  int sw_pair[NSLOT];
  real sw_lr[NSLOT];
  int sw_o = (p.iter_first - 1) % K;                  // offset of `it` inside its block of K iterations
  int sw_blk0 = (p.iter_first - 1) - sw_o;            // (first iteration of the block) - 1
  bool sw_fill = true;
  // recording countdown: iterations left until mod(it, n_interval) == 1
  int rec_left = ((1 - p.iter_first) % p.n_interval + p.n_interval) % p.n_interval;
  if (p.n_interval == 1) rec_left = -1;  // mod(it, 1) == 1 never holds (reference quirk Q6)

  mbar_wait(bar, 0);
  if constexpr (kF32) {
    // expand the staged tables into the packed station-pair records of htm_forward.cuh (once per launch)
    float4* s_x = reinterpret_cast<float4*>(s_obs + S);
    const float4* st = reinterpret_cast<const float4*>(s_sta);
    const float4* ob = reinterpret_cast<const float4*>(s_obs);
    for (int m = lane; m < (S + 1) / 2; m += 32) {
      const int j0 = 2 * m, j1 = j0 + 1;
      const StaRecF a = expand_station(st[j0], ob[j0], pxy.x, pxy.y);
      StaRecF b = a;
      if (j1 < S) {
        b = expand_station(st[j1], ob[j1], pxy.x, pxy.y);
      } else {  // odd tail: zero-weight copy
        b.B = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      store_station_pair(s_x + 4 * m, a, b);
    }
    __syncwarp();
  }
  // float32: each chain carries the (negated) weighted mean residuals of its accepted state as the shift
  float nct[NSLOT], nca[NSLOT];
  if constexpr (kF32) {
    const float4* s_x = reinterpret_cast<const float4*>(s_obs + S);
    float hx[NSLOT], hy[NSLOT], hz[NSLOT], z0[NSLOT], S1t[NSLOT], S1a[NSLOT], S2[NSLOT];
#pragma unroll
    for (int q = 0; q < NSLOT; ++q) {
      hx[q] = x[q] - pxy.x;
      hy[q] = y[q] - pxy.y;
      hz[q] = z[q];
      z0[q] = 0.f;
    }
    forward_pairs<NSLOT>(s_x, (S + 1) / 2, hx, hy, hz, g, z0, z0, S1t, S1a, S2);
#pragma unroll
    for (int q = 0; q < NSLOT; ++q) {
      nct[q] = -S1t[q] * evc.y;
      nca[q] = -S1a[q] * evc.z;
    }
  }

  for (int it = p.iter_first; it <= p.iter_last; ++it) {
    // ---- propose (all slots) ----
    int icmp[NSLOT];
    real nx[NSLOT], ny[NSLOT], nz[NSLOT], lpr[NSLOT], nlgz[NSLOT];
    bool ok[NSLOT];
    uint32_t wacc[NSLOT];
#pragma unroll
    for (int q = 0; q < NSLOT; ++q) {
      const uint32_t gid = (eg * R + rr[q]) * K + k;
      const u32x4 w = philox4x32_10(p.rk, static_cast<uint32_t>(it), gid, PHX_STEP, 0u);
      wacc[q] = w.v[3];
      // model_perturb (src/cls_model.f90:162-190) for component icmp: 0 -> z, 1 -> y, 2 -> x
      const int ic = static_cast<int>(below(w.v[0], 3u));
      icmp[q] = ic;
      const real gs = M<real>::gauss(w.v[1], w.v[2]);
      const bool isz = ic == 0;
      const real x_old = isz ? z[q] : (ic == 1 ? y[q] : x[q]);
      const real mu = isz ? p.prior_z : (ic == 1 ? pxy.y : pxy.x);
      const real step = isz ? p.step_z : p.step_xy;
      const real x_new = x_old + gs * step;
      const real dn = x_new - mu, dl = x_old - mu;
      real lp = -(dn * dn - dl * dl) * (isz ? p.inv2s2_z : p.inv2s2_xy);
      // type-1 prior: + ln(x_new - mu) - ln(x_old - mu); the second log is carried in lgz
      const real lgn = M<real>::log(isz ? fabs(dn) : static_cast<real>(1));
      nlgz[q] = isz ? lgn : lgz[q];
      lp = isz ? (lp + lgn - lgz[q]) : lp;
      ok[q] = !isz || (x_new > mu);
      lpr[q] = lp;
      nx[q] = ic == 2 ? x_new : x[q];
      ny[q] = ic == 1 ? x_new : y[q];
      nz[q] = isz ? x_new : z[q];
    }
    // ---- forward: one pass over the stations serves all NSLOT chains of this thread ----
    real S1t[NSLOT], S1a[NSLOT], S2[NSLOT];
    if constexpr (kF32) {
      const float4* s_x = reinterpret_cast<const float4*>(s_obs + S);
      float hx[NSLOT], hy[NSLOT], hz[NSLOT];
#pragma unroll
      for (int q = 0; q < NSLOT; ++q) {
        hx[q] = nx[q] - pxy.x;
        hy[q] = ny[q] - pxy.y;
        hz[q] = nz[q];
      }
      forward_pairs<NSLOT>(s_x, (S + 1) / 2, hx, hy, hz, g, nct, nca, S1t, S1a, S2);
    } else {
      real nct64[NSLOT], nca64[NSLOT];
      const real4 st = s_sta[0];
      const real4 ob = s_obs[0];
#pragma unroll
      for (int q = 0; q < NSLOT; ++q) {
        real ct, ca;
        station_resid(nx[q], ny[q], nz[q], g, st, ob, static_cast<real>(0), static_cast<real>(0), ct, ca);
        nct64[q] = -ct;
        nca64[q] = -ca;
        S1t[q] = 0;
        S1a[q] = 0;
        S2[q] = 0;
      }
#pragma unroll kScalarUnroll
      for (int j = 1; j < S; ++j) {
        const real4 stj = s_sta[j];
        const real4 obj = s_obs[j];
#pragma unroll
        for (int q = 0; q < NSLOT; ++q)
          station_accum(nx[q], ny[q], nz[q], g, nct64[q], nca64[q], stj, obj, S1t[q], S1a[q], S2[q]);
      }
    }
    const bool rec_now = rec_left == 0;
    // ---- judge, count, record, swap: each phase for all slots, so that the warp-uniform branches (record
    //      iteration? swap draws to refill?) are taken once per iteration and the shuffles of the slots overlap ----
    bool cold[NSLOT];
#pragma unroll
    for (int q = 0; q < NSLOT; ++q) {
      const real Lnew = finish_loglik<real>(S1t[q], S2[q], S1a[q], static_cast<real>(0), evc);
      // mcmc_judge_model (src/cls_mcmc.f90:193-203)
      const real ratio = M<real>::div(Lnew - L[q], T[q], iT[q]) + lpr[q];
      const real ru = M<real>::u_co(wacc[q]);
      const bool acc = ok[q] && (ru > static_cast<real>(0)) && (M<real>::log(ru) <= ratio);
      cold[q] = is_cold<real>(T[q]);
      {  // cold-chain counters, three 10-bit fields per word (emptied under the refill branch below, before a field can overflow);
         // invalid lanes clone a valid chain and must not count
        const uint32_t inc = (cold[q] && valid[q]) ? (1u << (10 * icmp[q])) : 0u;
        pk_p += inc;
        pk_a += acc ? inc : 0u;
      }
      x[q] = acc ? nx[q] : x[q];
      y[q] = acc ? ny[q] : y[q];
      z[q] = acc ? nz[q] : z[q];
      L[q] = acc ? Lnew : L[q];
      lgz[q] = acc ? nlgz[q] : lgz[q];
      if constexpr (kF32) {  // new shift = weighted mean residual of the accepted state
        nct[q] = acc ? fmaf(-static_cast<float>(S1t[q]), static_cast<float>(evc.y), nct[q]) : nct[q];
        nca[q] = acc ? fmaf(-static_cast<float>(S1a[q]), static_cast<float>(evc.z), nca[q]) : nca[q];
      }
      if (TRACE) {
        if (valid[q] && p.trace) {
          htm_step_trace t;
          t.proposal_type = 5 + icmp[q];
          t.index = 3 * (e + 1) - icmp[q];
          t.prior_ok = ok[q] ? 1 : 0;
          t.accepted = acc ? 1 : 0;
          t.log_likelihood = static_cast<double>(L[q]);
          p.trace[static_cast<size_t>(it - p.iter_first) * p.E * R * K + ci[q]] = t;
        }
      }
    }
    // record (src/hypo_tremor_mcmc.f90:270-280)
    if (rec_now) {
      const int slot = (it - 1) / p.n_interval - p.rec_origin;
#pragma unroll
      for (int q = 0; q < NSLOT; ++q) {
        const uint32_t coldmask = __ballot_sync(0xffffffffu, cold[q] && valid[q]);
        if (cold[q] && valid[q]) {
          const int m = __popc(coldmask & gmask & below_me);
          if (p.samples && slot >= 0 && slot < p.rec_cap && m < p.n_cool) {
            real4 rec;
            rec.x = x[q];
            rec.y = y[q];
            rec.z = z[q];
            rec.w = L[q];
            p.samples[((static_cast<size_t>(slot) * R + rr[q]) * p.n_cool + m) * p.E + e] = rec;
          }
          if (p.hist && it > p.n_burn) hist_add<real>(p, e, x[q], y[q], z[q], pxy.x, pxy.y);
        }
      }
    }
    // swap inside the tempering group (src/cls_parallel.f90:100-216, 285-302).  A group of ONE chain draws the pair
    // (0, 0) and exchanges its temperature with itself, so only the traced kernels (which write a swap record) need
    // the test.
    if (sw_fill) {  // first iteration of a block of K: the rare per-block work shares this one branch
      // the packed counters gain at most K * NSLOT per field until the next block starts: empty them before a
      // 10-bit field can overflow
      if ((pk_it + K) * NSLOT > 1023) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          cnt_p[c] += (pk_p >> (10 * c)) & 1023u;
          cnt_a[c] += (pk_a >> (10 * c)) & 1023u;
        }
        pk_p = 0;
        pk_a = 0;
        pk_it = 0;
      }
      pk_it += K;
    }
    if (TRACE ? K >= 2 : true) {
      if (sw_fill) {
#pragma unroll
        for (int q = 0; q < NSLOT; ++q) {
          const uint32_t grp = eg * R + rr[q];
          const u32x4 w = philox4x32_10(p.rk, static_cast<uint32_t>(sw_blk0 + k + 1), grp, PHX_SWAP, 0u);
          const int i1 = static_cast<int>(below(w.v[0], static_cast<uint32_t>(K)));
          int i2 = i1 + 1 + static_cast<int>(below(w.v[1], static_cast<uint32_t>(K - 1)));
          if (i2 >= K) i2 -= K;
          sw_pair[q] = i1 | (i2 << 8);
          const real ru2 = M<real>::u_co(w.v[2]);
          // ln r; r = 0 can never be accepted (r >= eps fails, :295): use -inf surrogate
          sw_lr[q] = ru2 > static_cast<real>(0) ? M<real>::log(ru2) : static_cast<real>(3.0e38);
        }
      }
      int pair[NSLOT];
      real lr[NSLOT], Lp[NSLOT], Tp[NSLOT], iTp[NSLOT];
#pragma unroll
      for (int q = 0; q < NSLOT; ++q) {
        pair[q] = __shfl_sync(0xffffffffu, sw_pair[q], base + sw_o);
        lr[q] = __shfl_sync(0xffffffffu, sw_lr[q], base + sw_o);
      }
#pragma unroll
      for (int q = 0; q < NSLOT; ++q) {
        const int i1 = pair[q] & 0xff, i2 = pair[q] >> 8;
        // the two chains of the pair read each other; every other lane reads garbage it ignores
        const int partner = (k == i1) ? i2 : i1;
        Lp[q] = __shfl_sync(0xffffffffu, L[q], base + partner);
        Tp[q] = __shfl_sync(0xffffffffu, T[q], base + partner);
        iTp[q] = __shfl_sync(0xffffffffu, iT[q], base + partner);
      }
#pragma unroll
      for (int q = 0; q < NSLOT; ++q) {
        const int i1 = pair[q] & 0xff, i2 = pair[q] >> 8;
        // del_s = (L2 - L1)(1/T1 - 1/T2) is symmetric under exchanging the roles of 1 and 2
        const real del_s = (Lp[q] - L[q]) * (iT[q] - iTp[q]);
        const bool sacc = lr[q] <= del_s;
        const bool mine = (k == i1) || (k == i2);
        if (TRACE) {
          const int a1 = __shfl_sync(0xffffffffu, sacc ? 1 : 0, base + i1);
          if (valid[q] && k == 0 && p.swaps) {
            htm_swap_trace t;
            t.rank1 = rr[q];
            t.chain1 = i1 + 1;
            t.rank2 = rr[q];
            t.chain2 = i2 + 1;
            t.accepted = a1;
            t.reserved = 0;
            p.swaps[(static_cast<size_t>(it - p.iter_first) * p.E + e) * R + rr[q]] = t;
          }
        }
        if (sacc && mine) {
          T[q] = Tp[q];
          iT[q] = iTp[q];
        }
      }
    }
    // advance the swap-draw window and the recording countdown
    sw_fill = false;
    if (++sw_o == K) {
      sw_o = 0;
      sw_blk0 += K;
      sw_fill = true;
    }
    rec_left = rec_now ? p.n_interval - 1 : rec_left - 1;
  }

  // ---- write back ----
#pragma unroll
  for (int q = 0; q < NSLOT; ++q) {
    if (valid[q]) {
      p.x[ci[q]] = x[q];
      p.y[ci[q]] = y[q];
      p.z[ci[q]] = z[q];
      p.L[ci[q]] = L[q];
      p.T[ci[q]] = T[q];
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {  // proposal types 5,6,7 = icmp 0,1,2 (src/cls_mcmc.f90:163)
    cnt_p[c] += (pk_p >> (10 * c)) & 1023u;
    cnt_a[c] += (pk_a >> (10 * c)) & 1023u;
    const uint32_t sp = warp_sum<uint32_t>(cnt_p[c]);
    const uint32_t sa = warp_sum<uint32_t>(cnt_a[c]);
    if (lane == 0 && p.counts) {
      atomicAdd(p.counts + 4 + c, static_cast<unsigned long long>(sp));
      atomicAdd(p.counts + 7 + 4 + c, static_cast<unsigned long long>(sa));
    }
  }
}

// ================================================================================================
// Wide-group kernel: tempering groups of MORE than 32 chains (the reference puts no limit on n_chains,
// src/hypo_tremor_mcmc.f90:114-118).  CTA = one group (event, rank): chain k sits in thread k of ceil(K/32) warps.
// Same per-chain step as the lane kernel with one chain per lane; what a warp shuffle did there -- the swap
// partner's (L, T), the pair and ln r of the iteration's swap attempt, the numbering of the cold chains for the
// records -- goes through double-buffered shared memory with ONE block barrier per iteration.
// ================================================================================================
template <typename real, bool TRACE>
__global__ void __launch_bounds__(1024) fact_wide_kernel(const FactParams<real> p) {
  typedef typename M<real>::real4 real4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr bool kF32 = sizeof(real) == 4;
  const int K = p.K, S = p.S, R = p.R, n_pairs = (S + 1) / 2;
  const int k = threadIdx.x, lane = threadIdx.x & 31;
  const int e = blockIdx.x / R, r = blockIdx.x % R;
  const bool valid = k < K;
  const int kk = valid ? k : K - 1;  // idle threads of the last warp clone a chain and never write
  // shared memory: staged tables | (float32) expanded pair records | barrier | published (L, T)[2][K] | swap draw [2]
  real4* s_sta = reinterpret_cast<real4*>(smem_raw);
  real4* s_obs = s_sta + S;
  float4* s_x = reinterpret_cast<float4*>(s_obs + S);
  unsigned char* after = reinterpret_cast<unsigned char*>(kF32 ? reinterpret_cast<real4*>(s_x + 4 * n_pairs) : s_obs + S);
  uint64_t* bar = reinterpret_cast<uint64_t*>(after);
  real* s_L = reinterpret_cast<real*>(bar + 2);  // [2][K]
  real* s_T = s_L + 2 * K;                       // [2][K]
  real* s_lr = s_T + 2 * K;                      // [2]
  int* s_pair = reinterpret_cast<int*>(s_lr + 2);  // [2]
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
    fence_proxy_async();
    const uint32_t bytes = static_cast<uint32_t>(S * sizeof(real4));
    mbar_expect_tx(bar, 2 * bytes);
    tma_load_1d(s_sta, p.sta4, bytes, bar);
    tma_load_1d(s_obs, p.obs4 + static_cast<size_t>(e) * S, bytes, bar);
  }
  __syncthreads();
  const real4 evc = p.evc4[e];
  const typename R2<real>::type pxy = p.prior_xy[e];
  const Glob<real> g = make_glob<real>(p.vs, p.qs);
  const uint32_t eg = static_cast<uint32_t>(e) + p.event_offset;
  const size_t ci = (static_cast<size_t>(e) * R + r) * K + kk;
  const uint32_t gid = (eg * R + r) * K + kk, grp = eg * R + r;
  real x = p.x[ci], y = p.y[ci], z = p.z[ci], L = p.L[ci], T = p.T[ci];
  uint32_t cnt_p[3] = {0, 0, 0}, cnt_a[3] = {0, 0, 0};
  mbar_wait(bar, 0);
  if constexpr (kF32) {  // expand the staged tables into packed station-pair records (htm_forward.cuh), once
    const float4* st = reinterpret_cast<const float4*>(s_sta);
    const float4* ob = reinterpret_cast<const float4*>(s_obs);
    for (int m = threadIdx.x; m < n_pairs; m += blockDim.x) {
      const int j0 = 2 * m, j1 = j0 + 1;
      const StaRecF a = expand_station(st[j0], ob[j0], pxy.x, pxy.y);
      StaRecF b = a;
      if (j1 < S)
        b = expand_station(st[j1], ob[j1], pxy.x, pxy.y);
      else
        b.B = make_float4(0.f, 0.f, 0.f, 0.f);
      store_station_pair(s_x + 4 * m, a, b);
    }
    __syncthreads();
  }
  // float32: the chain carries the (negated) weighted mean residuals of its accepted state as the shift
  float nct[1] = {0.f}, nca[1] = {0.f};
  if constexpr (kF32) {
    const float hx[1] = {x - pxy.x}, hy[1] = {y - pxy.y}, hz[1] = {z}, z0[1] = {0.f};
    float a1t[1], a1a[1], a2[1];
    forward_pairs<1>(s_x, n_pairs, hx, hy, hz, g, z0, z0, a1t, a1a, a2);
    nct[0] = -a1t[0] * evc.y;
    nca[0] = -a1a[0] * evc.z;
  }
  for (int it = p.iter_first; it <= p.iter_last; ++it) {
    const int buf = it & 1;
    const u32x4 w = philox4x32_10(p.rk, static_cast<uint32_t>(it), gid, PHX_STEP, 0u);
    int icmp;
    real nx, ny, nz, lpr;
    bool ok;
    propose_hypo<real>(w, x, y, z, pxy.x, pxy.y, p, icmp, nx, ny, nz, lpr, ok);
    real S1t, S1a, S2;
    if constexpr (kF32) {
      const float hx[1] = {nx - pxy.x}, hy[1] = {ny - pxy.y}, hz[1] = {nz};
      float a1t[1], a1a[1], a2[1];
      forward_pairs<1>(s_x, n_pairs, hx, hy, hz, g, nct, nca, a1t, a1a, a2);
      S1t = a1t[0];
      S1a = a1a[0];
      S2 = a2[0];
    } else {
      real ct, ca;
      station_resid(nx, ny, nz, g, s_sta[0], s_obs[0], static_cast<real>(0), static_cast<real>(0), ct, ca);
      S1t = S1a = S2 = 0;
      for (int j = 1; j < S; ++j) station_accum(nx, ny, nz, g, -ct, -ca, s_sta[j], s_obs[j], S1t, S1a, S2);
    }
    const real Lnew = finish_loglik<real>(S1t, S2, S1a, static_cast<real>(0), evc);
    const bool acc = judge<real>(Lnew, L, T, lpr, ok, w.v[3]);
    const bool cold = is_cold<real>(T);
    if (cold && valid) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        cnt_p[c] += (icmp == c) ? 1u : 0u;
        cnt_a[c] += (acc && icmp == c) ? 1u : 0u;
      }
    }
    if (acc) {
      x = nx;
      y = ny;
      z = nz;
      L = Lnew;
      if constexpr (kF32) {
        nct[0] = fmaf(-static_cast<float>(S1t), static_cast<float>(evc.y), nct[0]);
        nca[0] = fmaf(-static_cast<float>(S1a), static_cast<float>(evc.z), nca[0]);
      }
    }
    if (TRACE) {
      if (valid && p.trace) {
        htm_step_trace t;
        t.proposal_type = 5 + icmp;
        t.index = 3 * (e + 1) - icmp;
        t.prior_ok = ok ? 1 : 0;
        t.accepted = acc ? 1 : 0;
        t.log_likelihood = static_cast<double>(L);
        p.trace[static_cast<size_t>(it - p.iter_first) * p.E * R * K + ci] = t;
      }
    }
    // publish (L, T) and the iteration's swap attempt; one barrier; then everyone reads
    if (valid) {
      s_L[buf * K + k] = L;
      s_T[buf * K + k] = T;
    }
    if (k == 0) {
      const u32x4 ws = philox4x32_10(p.rk, static_cast<uint32_t>(it), grp, PHX_SWAP, 0u);
      const int i1 = static_cast<int>(below(ws.v[0], static_cast<uint32_t>(K)));
      int i2 = i1 + 1 + static_cast<int>(below(ws.v[1], static_cast<uint32_t>(K - 1)));
      if (i2 >= K) i2 -= K;
      const real ru = M<real>::u_co(ws.v[2]);
      s_pair[buf] = i1 | (i2 << 16);
      s_lr[buf] = ru > static_cast<real>(0) ? M<real>::log(ru) : static_cast<real>(3.0e38);  // r = 0 never accepts
    }
    __syncthreads();
    const real* bL = s_L + buf * K;
    const real* bT = s_T + buf * K;
    // record (src/hypo_tremor_mcmc.f90:270-280): slot m = this chain's rank among the group's cold chains
    if (p.n_interval > 1 && (it % p.n_interval) == 1 && cold && valid) {
      int m = 0;
      for (int q = 0; q < k; ++q) m += is_cold<real>(bT[q]) ? 1 : 0;
      const int slot = (it - 1) / p.n_interval - p.rec_origin;
      if (p.samples && slot >= 0 && slot < p.rec_cap && m < p.n_cool) {
        real4 rec;
        rec.x = x;
        rec.y = y;
        rec.z = z;
        rec.w = L;
        p.samples[((static_cast<size_t>(slot) * R + r) * p.n_cool + m) * p.E + e] = rec;
      }
      if (p.hist && it > p.n_burn) hist_add<real>(p, e, x, y, z, pxy.x, pxy.y);
    }
    // swap inside the group (src/cls_parallel.f90:100-216, 285-302): temperatures are exchanged
    {
      const int i1 = s_pair[buf] & 0xffff, i2 = s_pair[buf] >> 16;
      const real L1 = bL[i1], L2 = bL[i2], T1 = bT[i1], T2 = bT[i2];
      const real del_s = (L2 - L1) * (static_cast<real>(1) / T1 - static_cast<real>(1) / T2);
      const bool sacc = s_lr[buf] <= del_s;
      if (sacc && valid) {
        if (k == i1)
          T = T2;
        else if (k == i2)
          T = T1;
      }
      if (TRACE) {
        if (k == 0 && p.swaps) {
          htm_swap_trace t;
          t.rank1 = r;
          t.chain1 = i1 + 1;
          t.rank2 = r;
          t.chain2 = i2 + 1;
          t.accepted = sacc ? 1 : 0;
          t.reserved = 0;
          p.swaps[(static_cast<size_t>(it - p.iter_first) * p.E + e) * R + r] = t;
        }
      }
    }
  }
  if (valid) {
    p.x[ci] = x;
    p.y[ci] = y;
    p.z[ci] = z;
    p.L[ci] = L;
    p.T[ci] = T;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const uint32_t sp = warp_sum<uint32_t>(cnt_p[c]), sa = warp_sum<uint32_t>(cnt_a[c]);
    if (lane == 0 && p.counts) {
      if (sp) atomicAdd(p.counts + 4 + c, static_cast<unsigned long long>(sp));
      if (sa) atomicAdd(p.counts + 7 + 4 + c, static_cast<unsigned long long>(sa));
    }
  }
}

// ================================================================================================
// Warp-per-chain kernel.  CTA = gpc tempering groups x K warps, all of ONE event.
// ================================================================================================
template <typename real, int SPL, bool TRACE>
__global__ void __launch_bounds__(512) fact_warp_kernel(const FactParams<real> p, const int gpc) {
  typedef typename M<real>::real4 real4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = p.K, S = p.S, R = p.R;
  const int cpe = (R + gpc - 1) / gpc;  // CTAs per event
  const int e = blockIdx.x / cpe;
  const int gl = warp / K, k = warp % K;
  const int r = (blockIdx.x % cpe) * gpc + gl;

  real4* s_sta = reinterpret_cast<real4*>(smem_raw);
  real4* s_obs = s_sta + S;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(2) * S * sizeof(real4));
  real* s_L = reinterpret_cast<real*>(bar + 2);  // [2][gpc*K]
  real* s_T = s_L + 2 * gpc * K;                 // [2][gpc*K]
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
    fence_proxy_async();
    const uint32_t bytes = static_cast<uint32_t>(S * sizeof(real4));
    mbar_expect_tx(bar, 2 * bytes);
    tma_load_1d(s_sta, p.sta4, bytes, bar);
    tma_load_1d(s_obs, p.obs4 + static_cast<size_t>(e) * S, bytes, bar);
  }
  __syncthreads();  // barrier init visible to every waiter
  mbar_wait(bar, 0);
  if (r >= R) return;  // a whole group leaves together; it shares no barrier with the others

  // this lane's stations live in registers for the whole launch
  real4 st[SPL], ob[SPL];
#pragma unroll
  for (int s = 0; s < SPL; ++s) {
    const int j = lane + 32 * s;
    if (j < S) {
      st[s] = s_sta[j];
      ob[s] = s_obs[j];
    } else {  // padding lane: zero weight, harmless geometry
      st[s] = s_sta[0];
      ob[s] = s_obs[0];
      ob[s].y = 0;
      ob[s].w = 0;
    }
  }
  const size_t ci = (static_cast<size_t>(e) * R + r) * K + k;
  real x = p.x[ci], y = p.y[ci], z = p.z[ci], L = p.L[ci], T = p.T[ci];
  const real4 evc = p.evc4[e];
  const typename R2<real>::type pxy = p.prior_xy[e];
  const Glob<real> g = make_glob<real>(p.vs, p.qs);
  const uint32_t eg = static_cast<uint32_t>(e) + p.event_offset;
  const uint32_t gid = (eg * R + r) * K + k;
  const uint32_t grp = eg * R + r;
  uint32_t cnt_p[3] = {0, 0, 0}, cnt_a[3] = {0, 0, 0};

  for (int it = p.iter_first; it <= p.iter_last; ++it) {
    const u32x4 w = philox4x32_10(p.seed, static_cast<uint32_t>(it), gid, PHX_STEP, 0u);
    int icmp;
    real nx, ny, nz, lpr;
    bool ok;
    propose_hypo<real>(w, x, y, z, pxy.x, pxy.y, p, icmp, nx, ny, nz, lpr, ok);
    // forward: stations across lanes, one pass around the shift of station 0
    real rt[SPL], ra[SPL];
#pragma unroll
    for (int s = 0; s < SPL; ++s)
      station_resid(nx, ny, nz, g, st[s], ob[s], static_cast<real>(0), static_cast<real>(0), rt[s], ra[s]);
    const real ct = __shfl_sync(0xffffffffu, rt[0], 0);
    const real ca = __shfl_sync(0xffffffffu, ra[0], 0);
    real S1t = 0, S1a = 0, S2 = 0;
#pragma unroll
    for (int s = 0; s < SPL; ++s) {
      const real et = rt[s] - ct, ea = ra[s] - ca;
      const real qt = ob[s].y * et, qa = ob[s].w * ea;
      S1t += qt;
      S1a += qa;
      S2 += qt * et;
      S2 += qa * ea;
    }
    S1t = warp_sum(S1t);
    S1a = warp_sum(S1a);
    S2 = warp_sum(S2);
    const real Lnew = finish_loglik<real>(S1t, S2, S1a, static_cast<real>(0), evc);
    const bool acc = judge<real>(Lnew, L, T, lpr, ok, w.v[3]);
    const bool cold = is_cold<real>(T);
    if (cold) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        cnt_p[c] += (icmp == c) ? 1u : 0u;
        cnt_a[c] += (acc && icmp == c) ? 1u : 0u;
      }
    }
    if (acc) {
      x = nx;
      y = ny;
      z = nz;
      L = Lnew;
    }
    if (TRACE) {
      if (lane == 0 && p.trace) {
        htm_step_trace t;
        t.proposal_type = 5 + icmp;
        t.index = 3 * (e + 1) - icmp;
        t.prior_ok = ok ? 1 : 0;
        t.accepted = acc ? 1 : 0;
        t.log_likelihood = static_cast<double>(L);
        p.trace[static_cast<size_t>(it - p.iter_first) * p.E * R * K + ci] = t;
      }
    }
    // publish (L, T) for the record rank and the swap
    const int buf = it & 1;
    real* bL = s_L + (buf * gpc + gl) * K;
    real* bT = s_T + (buf * gpc + gl) * K;
    if (K >= 2 || (it % p.n_interval) == 1) {
      if (lane == 0) {
        bL[k] = L;
        bT[k] = T;
      }
      // named barrier of this tempering group only (ids 1..15)
      asm volatile("bar.sync %0, %1;" ::"r"(1 + gl), "r"(K * 32) : "memory");
    }
    if ((it % p.n_interval) == 1 && cold) {
      int m = 0;
      for (int kk = 0; kk < k; ++kk) m += is_cold<real>(bT[kk]) ? 1 : 0;
      const int slot = (it - 1) / p.n_interval - p.rec_origin;
      if (lane == 0) {
        if (p.samples && slot >= 0 && slot < p.rec_cap && m < p.n_cool) {
          real4 rec;
          rec.x = x;
          rec.y = y;
          rec.z = z;
          rec.w = L;
          p.samples[((static_cast<size_t>(slot) * R + r) * p.n_cool + m) * p.E + e] = rec;
        }
        if (p.hist && it > p.n_burn) hist_add<real>(p, e, x, y, z, pxy.x, pxy.y);
      }
    }
    if (K >= 2) {
      const u32x4 ws = philox4x32_10(p.seed, static_cast<uint32_t>(it), grp, PHX_SWAP, 0u);
      const int i1 = static_cast<int>(below(ws.v[0], static_cast<uint32_t>(K)));
      int i2 = i1 + 1 + static_cast<int>(below(ws.v[1], static_cast<uint32_t>(K - 1)));
      if (i2 >= K) i2 -= K;
      const real L1 = bL[i1], L2 = bL[i2], T1 = bT[i1], T2 = bT[i2];
      const bool sacc = judge_swap<real>(T1, T2, L1, L2, ws.v[2]);
      if (sacc) {
        if (k == i1)
          T = T2;
        else if (k == i2)
          T = T1;
      }
      if (TRACE) {
        if (lane == 0 && k == 0 && p.swaps) {
          htm_swap_trace t;
          t.rank1 = r;
          t.chain1 = i1 + 1;
          t.rank2 = r;
          t.chain2 = i2 + 1;
          t.accepted = sacc ? 1 : 0;
          t.reserved = 0;
          p.swaps[(static_cast<size_t>(it - p.iter_first) * p.E + e) * R + r] = t;
        }
      }
    }
  }
  if (lane == 0) {
    p.x[ci] = x;
    p.y[ci] = y;
    p.z[ci] = z;
    p.L[ci] = L;
    p.T[ci] = T;
    if (p.counts) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (cnt_p[c]) atomicAdd(p.counts + 4 + c, static_cast<unsigned long long>(cnt_p[c]));
        if (cnt_a[c]) atomicAdd(p.counts + 7 + 4 + c, static_cast<unsigned long long>(cnt_a[c]));
      }
    }
  }
}

// ================================================================================================
// Chain set-up (replaces src/hypo_tremor_mcmc.f90:120-211 with Philox draws): one thread per chain
// ================================================================================================
template <typename real>
__global__ void fact_init_kernel(const FactParams<real> p, const real temp_high, const int ladder) {
  typedef typename M<real>::real4 real4;
  const size_t n = static_cast<size_t>(p.E) * p.R * p.K;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int k = static_cast<int>(i % p.K);
  const int e = static_cast<int>(i / (static_cast<size_t>(p.R) * p.K));
  const int r = static_cast<int>((i / p.K) % p.R);
  const uint32_t gid = ((static_cast<uint32_t>(e) + p.event_offset) * p.R + r) * p.K + k;
  const u32x4 a = philox4x32_10(p.seed, 0u, gid, PHX_INIT, 0u);
  const u32x4 b = philox4x32_10(p.seed, 1u, gid, PHX_INIT, 0u);
  const typename R2<real>::type pxy = p.prior_xy[e];
  // generate_model, src/cls_model.f90:139-158: Gaussian x, y; Rayleigh z - prior_z
  const real x = pxy.x + M<real>::gauss(a.v[0], a.v[1]) * p.width_xy;
  const real y = pxy.y + M<real>::gauss(a.v[2], a.v[3]) * p.width_xy;
  const real z = p.prior_z + M<real>::sqrt(static_cast<real>(-2) * M<real>::log(M<real>::u_oo(b.v[0]))) * p.width_z;
  real T = 1;
  if (k >= p.n_cool) {
    if (ladder == HTM_LADDER_GEOMETRIC) {
      const int n_hot = p.K - p.n_cool;
      T = M<real>::exp(M<real>::log(temp_high) * static_cast<real>(k - p.n_cool + 1) / static_cast<real>(n_hot));
    } else {  // src/hypo_tremor_mcmc.f90:205-206
      const u32x4 t = philox4x32_10(p.seed, 0u, gid, PHX_TEMP, 0u);
      const real u = M<real>::u_co(t.v[0]);
      T = M<real>::exp((u * (static_cast<real>(1) - static_cast<real>(kEps64)) + static_cast<real>(kEps64)) *
                       M<real>::log(temp_high));
    }
    // The reference keeps a hot chain off T = 1 by mixing in eps(1.d0) (src/hypo_tremor_mcmc.f90:205); in
    // float32 that margin rounds away (u = 0 gave T == 1.0f and one extra "cold" chain in a 6.4 M-chain
    // run), so the smallest float above 1 is enforced instead.
    if (sizeof(real) == 4 && !(T > static_cast<real>(1))) T = static_cast<real>(1.00000012f);
  }
  const Glob<real> g = make_glob<real>(p.vs, p.qs);
  const real L = lane_event_loglik<real>(p.sta4, p.obs4 + static_cast<size_t>(e) * p.S, p.evc4[e], p.S, x, y, z, g);
  p.x[i] = x;
  p.y[i] = y;
  p.z[i] = z;
  p.L[i] = L;
  p.T[i] = T;
}

// ================================================================================================
// Host launchers
// ================================================================================================
template <typename real>
static FactParams<real> make_params(const FactLaunch& a) {
  FactParams<real> p;
  typedef typename M<real>::real4 real4;
  p.sta4 = static_cast<const real4*>(a.tab.sta4);
  p.obs4 = static_cast<const real4*>(a.tab.obs4);
  p.evc4 = static_cast<const real4*>(a.tab.evc4);
  p.prior_xy = static_cast<const typename R2<real>::type*>(a.tab.prior_xy);
  p.x = static_cast<real*>(a.x);
  p.y = static_cast<real*>(a.y);
  p.z = static_cast<real*>(a.z);
  p.L = static_cast<real*>(a.L);
  p.T = static_cast<real*>(a.T);
  p.E = a.E;
  p.S = a.S;
  p.R = a.R;
  p.K = a.K;
  p.n_cool = a.n_cool;
  p.iter_first = a.iter_first;
  p.iter_last = a.iter_last;
  p.n_burn = a.n_burn;
  p.n_interval = a.n_interval;
  p.seed = a.seed;
  p.rk = philox_keys(a.seed);
  p.event_offset = a.event_offset;
  p.vs = static_cast<real>(a.vs);
  p.qs = static_cast<real>(a.qs);
  p.g_ivs = static_cast<real>(1) / p.vs;
  p.g_B = static_cast<real>(kPi * kFreq) / (p.qs * p.vs);
  p.prior_z = static_cast<real>(a.prior_z);
  p.width_z = static_cast<real>(a.width_z);
  p.width_xy = static_cast<real>(a.width_xy);
  p.step_xy = static_cast<real>(a.step_xy);
  p.step_z = static_cast<real>(a.step_z);
  p.inv2s2_xy = static_cast<real>(1.0 / (2.0 * a.width_xy * a.width_xy));
  p.inv2s2_z = static_cast<real>(1.0 / (2.0 * a.width_z * a.width_z));
  p.counts = a.counts;
  p.samples = static_cast<real4*>(a.samples);
  p.rec_origin = a.rec_origin;
  p.rec_cap = a.rec_cap;
  p.hist = a.hist;
  p.hist_bins = a.hist_bins;
  p.hist_hw = static_cast<real>(a.hist_hw);
  p.hist_zmax = static_cast<real>(a.hist_zmax);
  p.trace = a.trace;
  p.swaps = a.swaps;
  return p;
}

template <typename real, int NSLOT>
static cudaError_t launch_lane(const FactLaunch& a, cudaStream_t stream) {
  typedef typename M<real>::real4 real4;
  const FactParams<real> p = make_params<real>(a);
  const int gpw = a.R < 32 / a.K ? a.R : 32 / a.K;
  const int gps = gpw * NSLOT;
  const int wpe = (a.R + gps - 1) / gps;
  const long n_warps = static_cast<long>(a.E) * wpe;
  // warps are independent, so the CTA size only sets how evenly they spread over the SMs:
  // 1-warp CTAs while everything is resident at once (<= 32 CTAs/SM), else 2 or 4
  int wpb = n_warps <= 148L * 32 ? 1 : (n_warps <= 148L * 64 ? 2 : 4);
  const size_t per_warp = ((sizeof(real) == 4 ? 4 : 2) * a.S + (sizeof(real) == 4 ? 2 : 0)) * sizeof(real4) + sizeof(uint64_t);
  while (wpb > 1 && wpb * per_warp > 200 * 1024) wpb >>= 1;  // very large station counts: fewer warps per CTA
  const unsigned grid = static_cast<unsigned>((n_warps + wpb - 1) / wpb);
  const size_t smem = static_cast<size_t>(wpb) * per_warp;
  const bool trace = a.trace || a.swaps;
  cudaError_t err;
  if (trace) {
    err = cudaFuncSetAttribute(fact_lane_kernel<real, NSLOT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem));
    if (err != cudaSuccess) return err;
    fact_lane_kernel<real, NSLOT, true><<<grid, wpb * 32, smem, stream>>>(p);
  } else {
    err = cudaFuncSetAttribute(fact_lane_kernel<real, NSLOT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem));
    if (err != cudaSuccess) return err;
    fact_lane_kernel<real, NSLOT, false><<<grid, wpb * 32, smem, stream>>>(p);
  }
  return cudaGetLastError();
}

template <typename real, int SPL>
static cudaError_t launch_warp(const FactLaunch& a, cudaStream_t stream) {
  typedef typename M<real>::real4 real4;
  const FactParams<real> p = make_params<real>(a);
  int gpc = 16 / a.K;  // aim at <= 16 warps per CTA
  if (gpc < 1) gpc = 1;
  if (gpc > a.R) gpc = a.R;
  if (gpc > 15) gpc = 15;  // named barrier ids 1..15
  const int cpe = (a.R + gpc - 1) / gpc;
  const unsigned grid = static_cast<unsigned>(a.E) * cpe;
  const unsigned block = static_cast<unsigned>(gpc * a.K * 32);
  const size_t smem = 2 * a.S * sizeof(real4) + 2 * sizeof(uint64_t) + 4 * gpc * a.K * sizeof(real);
  const bool trace = a.trace || a.swaps;
  cudaError_t err;
  if (trace) {
    err = cudaFuncSetAttribute(fact_warp_kernel<real, SPL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem));
    if (err != cudaSuccess) return err;
    fact_warp_kernel<real, SPL, true><<<grid, block, smem, stream>>>(p, gpc);
  } else {
    err = cudaFuncSetAttribute(fact_warp_kernel<real, SPL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem));
    if (err != cudaSuccess) return err;
    fact_warp_kernel<real, SPL, false><<<grid, block, smem, stream>>>(p, gpc);
  }
  return cudaGetLastError();
}

template <typename real>
static cudaError_t launch_wide(const FactLaunch& a, cudaStream_t stream, const char** why) {
  typedef typename M<real>::real4 real4;
  if (a.K > 1024) {
    *why = "factorised mode supports at most 1024 chains per tempering group (one CTA per group)";
    return cudaErrorInvalidValue;
  }
  const FactParams<real> p = make_params<real>(a);
  const int n_pairs = (a.S + 1) / 2;
  const size_t smem = 2 * a.S * sizeof(real4) + (sizeof(real) == 4 ? 4 * n_pairs * sizeof(float4) : 0) + 2 * sizeof(uint64_t) +
                      (4 * a.K + 2) * sizeof(real) + 2 * sizeof(int) + 16;
  if (smem > 200 * 1024) {
    *why = "n_sta too large for the shared-memory staging of the wide-group kernel";
    return cudaErrorInvalidValue;
  }
  const unsigned grid = static_cast<unsigned>(a.E) * a.R, block = static_cast<unsigned>((a.K + 31) / 32 * 32);
  const bool trace = a.trace || a.swaps;
  cudaError_t err;
  if (trace) {
    err = cudaFuncSetAttribute(fact_wide_kernel<real, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (err != cudaSuccess) return err;
    fact_wide_kernel<real, true><<<grid, block, smem, stream>>>(p);
  } else {
    err = cudaFuncSetAttribute(fact_wide_kernel<real, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (err != cudaSuccess) return err;
    fact_wide_kernel<real, false><<<grid, block, smem, stream>>>(p);
  }
  return cudaGetLastError();
}

template <typename real>
static cudaError_t launch_factorised_t(const FactLaunch& a, cudaStream_t stream, const char** why) {
  if (a.K > 32) {  // a tempering group wider than a warp: one CTA per group
    if (a.kernel == HTM_KERNEL_WARP_PER_CHAIN) {
      *why = "warp-per-chain kernel supports n_chains <= 16 (one CTA of <= 512 threads per tempering group)";
      return cudaErrorInvalidValue;
    }
    return launch_wide<real>(a, stream, why);
  }
  if (a.kernel == HTM_KERNEL_WARP_PER_CHAIN) {
    if (a.K > 16) {
      *why = "warp-per-chain kernel supports n_chains <= 16 (one CTA of <= 512 threads per tempering group)";
      return cudaErrorInvalidValue;
    }
    if (a.S <= 32) return launch_warp<real, 1>(a, stream);
    if (a.S <= 64) return launch_warp<real, 2>(a, stream);
    if (a.S <= 128) return launch_warp<real, 4>(a, stream);
    *why = "warp-per-chain kernel supports n_sta <= 128; use the lane-per-chain kernel";
    return cudaErrorInvalidValue;
  }
  const size_t smem_need = ((sizeof(real) == 4 ? 4 : 2) * a.S + 2) * sizeof(typename M<real>::real4) + 8;
  if (smem_need > 200 * 1024) {
    *why = "n_sta too large for the shared-memory staging of the lane-per-chain kernel";
    return cudaErrorInvalidValue;
  }
  // Chains per lane.  One keeps the most warps in flight; two share the station-pair reads and the loop
  // overhead between two chains and measure 6-7 % faster (float32) once the tempering groups fill both slot
  // rows exactly and the halved number of warps still oversubscribes the 592 schedulers about tenfold
  // (10 000 x 50 x 4 x 16: 2.76 -> 2.94e10 proposals/s; 1000 events: equal; 3000 events: 2 % slower).  Four
  // never won (tools/slots_sweep.py).
  int slots = a.slots;
  if (slots == 0) {
    const int gpw = a.R < 32 / a.K ? a.R : 32 / a.K;  // tempering groups side by side in one slot row
    const long warps2 = static_cast<long>(a.E) * (a.R / (2 * gpw));
    slots = (sizeof(real) == 4 && a.R % (2 * gpw) == 0 && warps2 >= 6000) ? 2 : 1;
  }
  switch (slots) {
    case 1: return launch_lane<real, 1>(a, stream);
    case 2: return launch_lane<real, 2>(a, stream);
    case 4: return launch_lane<real, 4>(a, stream);
    default: *why = "slots must be 1, 2 or 4"; return cudaErrorInvalidValue;
  }
}

cudaError_t launch_factorised(const FactLaunch& a, cudaStream_t stream, int* n_launches, const char** why) {
  static const char* none = "";
  *why = none;
  cudaError_t err = a.precision == HTM_PRECISION_F64 ? launch_factorised_t<double>(a, stream, why)
                                                     : launch_factorised_t<float>(a, stream, why);
  if (n_launches) *n_launches = err == cudaSuccess ? 1 : 0;
  return err;
}

cudaError_t launch_factorised_init(const FactLaunch& a, double temp_high, int ladder, cudaStream_t stream) {
  const size_t n = static_cast<size_t>(a.E) * a.R * a.K;
  const unsigned block = 128, grid = static_cast<unsigned>((n + block - 1) / block);
  if (a.precision == HTM_PRECISION_F64) {
    fact_init_kernel<double><<<grid, block, 0, stream>>>(make_params<double>(a), temp_high, ladder);
  } else {
    fact_init_kernel<float><<<grid, block, 0, stream>>>(make_params<float>(a), static_cast<float>(temp_high), ladder);
  }
  return cudaGetLastError();
}

}  // namespace htm
