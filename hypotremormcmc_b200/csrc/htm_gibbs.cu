// Mode C (blocked Gibbs): joint chains with the shared parameters solved.
//
// J = n_procs*n_chains joint chains, each = {vs, qs, t_corr(S), a_corr(S), T, all E hypocentres}.
// One iteration (the CPU statement of this exact schedule lives with the test oracle):
//   1. gibbs_sweep_kernel   every (chain, event) proposes one hypocentre coordinate and judges it on
//                           the event's own log-likelihood with the chain's temperature
//                           (src/cls_mcmc.f90:159-165,193-203; src/cls_forward.f90:307-362), then
//                           evaluates the chain's pending shared-parameter proposal for that event and
//                           writes per-tile partial sums of L (current and proposed).
//   2. gibbs_decide_kernel  one CTA: adds the partial sums in a fixed order, judges the shared-parameter
//                           proposal on the sum over ALL events (src/cls_forward.f90:268-303 is what the
//                           reference recomputes for such a move), records the cold chains' shared
//                           parameters, does the one swap attempt over all J chains
//                           (src/cls_parallel.f90:220-240,285-302) and draws the next proposal.
// A shared-parameter acceptance is committed lazily: the next sweep picks L_e := L_e(proposed).
//
// Mapping of the sweep: CTA = tile of 32 events x up to 8 chains; warp = chain, lane = event.  The
// tile's observation rows are staged by 32 bulk-TMA copies (one per event) into padded shared-memory
// rows, so the per-lane 16-byte reads are bank-conflict free; station table and the chains' station
// terms are warp-broadcast reads.  float32 at large E uses a second mapping (gibbs_sweep_oq_kernel:
// warp = 8 events x 4 chains, CTA walks event octets through a 2-stage TMA ring); the persistent
// cooperative kernel takes over whenever every tile can be resident at once.
#include <cooperative_groups.h>

#include <cstdlib>
#include <string>

#include "htm_gibbs_step.cuh"

namespace cg = cooperative_groups;

namespace htm {

__global__ void __launch_bounds__(256) gibbs_decide_kernel(const GibbsDecide d) {
  extern __shared__ __align__(16) unsigned char s_decide_dyn[];
  gibbs_decide(d, s_decide_dyn);
}


// builds the expanded rows once per table upload: one thread per (event, pair) and one per event header
__global__ void expand_obs_kernel(const float4* __restrict__ sta4, const float4* __restrict__ obs4,
                                  const float2* __restrict__ prior_xy, int E, int S, float4* __restrict__ obsx) {
  const int n_pairs = S / 2, per_ev = n_pairs + 1, xrow = 2 + 4 * n_pairs;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(E) * per_ev) return;
  const int e = static_cast<int>(i / per_ev), m = static_cast<int>(i % per_ev) - 1;
  const float2 c = prior_xy[e];
  const float4* ob = obs4 + static_cast<size_t>(e) * S;
  float4* row = obsx + static_cast<size_t>(e) * xrow;
  if (m < 0) {
    const StaRecF r0 = expand_station(sta4[0], ob[0], c.x, c.y);
    row[0] = r0.A;
    row[1] = make_float4(ob[0].x, ob[0].z, 0.f, 0.f);
    return;
  }
  const int j0 = 1 + 2 * m, j1 = j0 + 1;
  const StaRecF a = expand_station(sta4[j0], ob[j0], c.x, c.y);
  StaRecF b = a;
  if (j1 < S)
    b = expand_station(sta4[j1], ob[j1], c.x, c.y);
  else
    b.B = make_float4(0.f, 0.f, 0.f, 0.f);
  store_station_pair(row + 2 + 4 * m, a, b);
}

cudaError_t launch_expand_obs(const Tables& tab, int E, int S, void* obsx, cudaStream_t stream) {
  const size_t n = static_cast<size_t>(E) * (S / 2 + 1);
  expand_obs_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, stream>>>(
      static_cast<const float4*>(tab.sta4), static_cast<const float4*>(tab.obs4_raw),
      static_cast<const float2*>(tab.prior_xy), E, S, static_cast<float4*>(obsx));
  return cudaGetLastError();
}

// shared-memory carve-up common to both kernels
template <typename real>
struct SweepSm {
  typedef typename M<real>::real4 real4;
  uint64_t* bar;
  real4* obs;   // [kTile][row]
  real4* sta;   // f64: [S]
  real* tc;     // f64: [kCW][S]
  real* ac;     // f64: [kCW][S]
  float4* cp;   // f32: [kCW][n_pairs]
  float4* cpP;  // f32: [kCW][n_pairs]
  float4* c0;   // f32: [kCW]
  int row;
  unsigned char* end;  // first byte after the sweep's own shared memory (16-byte aligned)
};
template <typename real>
__device__ __forceinline__ SweepSm<real> carve_sweep_sm(unsigned char* base, int S, int xrow) {
  typedef typename M<real>::real4 real4;
  constexpr bool kF32 = sizeof(real) == 4;
  SweepSm<real> m;
  const int n_pairs = S / 2;
  m.row = kF32 ? xrow + 1 : S + 1;
  m.bar = reinterpret_cast<uint64_t*>(base);
  m.obs = reinterpret_cast<real4*>(base + 16);
  m.sta = m.obs + kTile * m.row;
  m.tc = reinterpret_cast<real*>(m.sta + S);
  m.ac = m.tc + kCW * S;
  m.cp = reinterpret_cast<float4*>(m.obs + kTile * m.row);
  m.cpP = m.cp + kCW * n_pairs;
  m.c0 = m.cpP + kCW * n_pairs;
  unsigned char* e = kF32 ? reinterpret_cast<unsigned char*>(m.c0 + kCW) : reinterpret_cast<unsigned char*>(m.ac + kCW * S);
  m.end = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(e) + 15) & ~static_cast<uintptr_t>(15));
  return m;
}
template <typename real>
static size_t sweep_smem(int S) {
  typedef typename M<real>::real4 real4;
  if (sizeof(real) == 4) {
    const int n_pairs = S / 2, xrow = 2 + 4 * n_pairs;
    return 16 + static_cast<size_t>(kTile) * (xrow + 1) * sizeof(float4) + (2 * kCW * n_pairs + kCW) * sizeof(float4) + 16;
  }
  return 16 + static_cast<size_t>(kTile) * (S + 1) * sizeof(real4) + S * sizeof(real4) + 2 * kCW * S * sizeof(real) + 16;
}

// TMA-stage the tile's observation rows (and, float64, the station table); caller syncs and waits
template <typename real>
__device__ __forceinline__ void stage_tile(const GibbsParams<real>& p, const SweepSm<real>& m, int tile, int n_ev) {
  typedef typename M<real>::real4 real4;
  constexpr bool kF32 = sizeof(real) == 4;
  const int lane = threadIdx.x & 31;
  if (kF32) {
    const uint32_t bytes = static_cast<uint32_t>(p.xrow * sizeof(float4));
    if (lane == 0) mbar_expect_tx(m.bar, bytes * n_ev);
    __syncwarp();
    if (lane < n_ev)
      tma_load_1d(m.obs + lane * m.row, p.obsx + static_cast<size_t>(tile * kTile + lane) * p.xrow, bytes, m.bar);
  } else {
    const uint32_t bytes = static_cast<uint32_t>(p.S * sizeof(real4));
    if (lane == 0) mbar_expect_tx(m.bar, bytes * (n_ev + 1));
    __syncwarp();
    if (lane < n_ev)
      tma_load_1d(m.obs + lane * m.row, p.obs4 + static_cast<size_t>(tile * kTile + lane) * p.S, bytes, m.bar);
    if (lane == 0) tma_load_1d(m.sta, p.sta4, bytes, m.bar);
  }
}

// the warp's chain operands in shared memory, from (tc, ac) arrays of any addressable memory
template <typename real>
__device__ __forceinline__ void stage_chain_terms(const SweepSm<real>& m, int warp, int S, const double* gtc,
                                                  const double* gac, int wh, int pi, double pvd) {
  constexpr bool kF32 = sizeof(real) == 4;
  const int lane = threadIdx.x & 31, n_pairs = S / 2;
  if (kF32) {
    const float pv = static_cast<float>(pvd);
    for (int mm = lane; mm < n_pairs; mm += 32) {
      const int j0 = 1 + 2 * mm, j1 = j0 + 1;
      float4 cur = make_float4(-static_cast<float>(gtc[j0]), 0.f, -static_cast<float>(gac[j0]), 0.f);
      if (j1 < S) {
        cur.y = -static_cast<float>(gtc[j1]);
        cur.w = -static_cast<float>(gac[j1]);
      }
      float4 prp = cur;
      if (wh == 2 && pi == j0) prp.x = -pv;
      if (wh == 2 && pi == j1) prp.y = -pv;
      if (wh == 4 && pi == j0) prp.z = -pv;
      if (wh == 4 && pi == j1) prp.w = -pv;
      m.cp[warp * n_pairs + mm] = cur;
      m.cpP[warp * n_pairs + mm] = prp;
    }
    if (lane == 0) {
      float4 c0 = make_float4(-static_cast<float>(gtc[0]), -static_cast<float>(gac[0]), 0.f, 0.f);
      c0.z = (wh == 2 && pi == 0) ? -pv : c0.x;
      c0.w = (wh == 4 && pi == 0) ? -pv : c0.y;
      m.c0[warp] = c0;
    }
  } else {
    for (int j = lane; j < S; j += 32) {
      m.tc[warp * S + j] = static_cast<real>(gtc[j]);
      m.ac[warp * S + j] = static_cast<real>(gac[j]);
    }
  }
}

template <typename real>
__device__ __forceinline__ StepIn<real> make_step_in(const SweepSm<real>& m, int warp, int S, double Td, double vs,
                                                     double qs, int which, int pidx, double pval,
                                                     const typename M<real>::real4* obs_row) {
  StepIn<real> in;
  in.T = static_cast<real>(Td);
  in.iT = static_cast<real>(1) / in.T;
  in.cold = gibbs_is_cold<real>(Td);
  in.vs = static_cast<real>(vs);
  in.qs = static_cast<real>(qs);
  in.which = which;
  in.pidx = pidx;
  in.pval = static_cast<real>(pval);
  in.S = S;
  in.n_pairs = S / 2;
  in.obs_row = obs_row;
  in.cp = m.cp + warp * in.n_pairs;
  in.cpP = m.cpP + warp * in.n_pairs;
  in.c0 = sizeof(real) == 4 ? m.c0[warp] : make_float4(0.f, 0.f, 0.f, 0.f);
  in.s_sta = m.sta;
  in.tc = m.tc + warp * S;
  in.ac = m.ac + warp * S;
  return in;
}

// per-tile partial sums (float64, butterfly = fixed order) and cold-chain counters
template <typename real>
__device__ __forceinline__ void tile_sums_and_counts(const GibbsParams<real>& p, double* part_cur, double* part_prop,
                                                     int c, int tile, bool ev_ok, bool cold, real Le, real Lp, int icmp,
                                                     bool acc) {
  const int lane = threadIdx.x & 31;
  const double s_cur = warp_sum<double>(ev_ok ? static_cast<double>(Le) : 0.0);
  const double s_prop = warp_sum<double>(ev_ok ? static_cast<double>(Lp) : 0.0);
  if (lane == 0) {
    part_cur[static_cast<size_t>(c) * p.n_tiles + tile] = s_cur;
    part_prop[static_cast<size_t>(c) * p.n_tiles + tile] = s_prop;
  }
  if (cold) {
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const uint32_t np = __popc(__ballot_sync(0xffffffffu, ev_ok && icmp == t));
      const uint32_t na = __popc(__ballot_sync(0xffffffffu, ev_ok && icmp == t && acc));
      if (lane == 0 && p.counts) {
        if (np) atomicAdd(p.counts + 4 + t, static_cast<unsigned long long>(np));
        if (na) atomicAdd(p.counts + 11 + t, static_cast<unsigned long long>(na));
      }
    }
  }
}

// Tail of both sweep kernels: the last CTA to finish judges the shared-parameter proposals (fixed-order sums:
// deterministic), reusing its shared memory for the chain-level state.  Every thread of the CTA must call it.
__device__ __forceinline__ void last_cta_decides(const GibbsDecide& dec, unsigned int* done_counter, unsigned char* smem) {
  if (done_counter == nullptr) return;  // event-sharded run: totals -> all-reduce -> decide are separate launches
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();  // the CTA's partial sums (ordered before by the barrier) become visible device-wide
    const unsigned int ticket = atomicAdd(done_counter, 1u);
    s_last = ticket == gridDim.x * gridDim.y - 1 ? 1 : 0;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    gibbs_decide(dec, smem);
    if (threadIdx.x == 0) *done_counter = 0u;
  }
}

// ---- one iteration per launch: sweep + (last CTA) decide --------------------------------------------------
template <typename real, bool TRACE>
__global__ void __launch_bounds__(kCW * 32, 3) gibbs_sweep_kernel(const GibbsParams<real> p, const GibbsDecide dec,
                                                                  unsigned int* done_counter) {
  typedef typename M<real>::real4 real4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.S, J = p.J, E = p.E;
  const int tile = blockIdx.x, e = tile * kTile + lane;
  const int c = blockIdx.y * kCW + warp;
  const SweepSm<real> m = carve_sweep_sm<real>(smem_raw, S, p.xrow);
  const int n_ev = min(kTile, E - tile * kTile);
  if (threadIdx.x == 0) {
    mbar_init(m.bar, 1);
    fence_mbar_init();
    fence_proxy_async();
  }
  __syncthreads();
  if (warp == 0) stage_tile<real>(p, m, tile, n_ev);
  const bool chain_ok = c < J;
  if (chain_ok)
    stage_chain_terms<real>(m, warp, S, p.g_tc + static_cast<size_t>(c) * S, p.g_ac + static_cast<size_t>(c) * S,
                            p.prop_which[c], p.prop_idx[c], p.prop_xnew[c]);
  __syncthreads();
  mbar_wait(m.bar, 0);
  if (chain_ok) {
    const bool ev_ok = e < E;
    const int ee = ev_ok ? e : E - 1;  // idle lanes clone the last event and never write
    const size_t ci = static_cast<size_t>(c) * E + ee;
    const StepIn<real> in = make_step_in<real>(m, warp, S, p.g_T[c], p.g_vs[c], p.g_qs[c], p.prop_which[c], p.prop_idx[c],
                                               p.prop_xnew[c], m.obs + (ev_ok ? lane : n_ev - 1) * m.row);
    const real4 evc = p.evc4[ee];
    const real mux = reinterpret_cast<const real*>(p.prior_xy)[2 * ee], muy = reinterpret_cast<const real*>(p.prior_xy)[2 * ee + 1];
    real x = p.hx[ci], y = p.hy[ci], z = p.hz[ci];
    real Le = p.a_prev[c] ? p.hLp[ci] : p.hLe[ci];  // lazy commit of the last shared-parameter acceptance
    real Lp;
    int icmp;
    bool acc;
    gibbs_thread_step<real, TRACE>(p, p.it, in, c, e, ee, ev_ok, evc, mux, muy, x, y, z, Le, Lp, icmp, acc, p.trace);
    if (ev_ok) {
      p.hx[ci] = x;
      p.hy[ci] = y;
      p.hz[ci] = z;
      p.hLe[ci] = Le;
      p.hLp[ci] = Lp;
      if (p.rec_slot >= 0 && p.slot_of[c] >= 0 && p.hypo_rec) {
        real4 rec;
        rec.x = x;
        rec.y = y;
        rec.z = z;
        rec.w = Le;
        p.hypo_rec[(static_cast<size_t>(p.rec_slot) * p.n_cool_total + p.slot_of[c]) * E + e] = rec;
      }
    }
    tile_sums_and_counts<real>(p, p.part_cur, p.part_prop, c, tile, ev_ok, in.cold, Le, Lp, icmp, acc);
  }

  last_cta_decides(dec, done_counter, smem_raw);
}

// ---- float32 sweep for large E: warp = 8 events x 4 chains, CTA loops over event octets --------------------
// Same schedule and the same arithmetic per (chain, event) as gibbs_sweep_kernel.  What changes is the mapping:
//   * lane = (event of the octet, chain of the quad), chain minor.  A 16-byte row read touches 8 distinct rows
//     instead of 32 and neighbouring lanes share them: 2 shared-memory wavefronts instead of 4 (the unblocked
//     sweep is bound by exactly that traffic); the chains' station terms are 4 distinct 16-byte words.
//   * a CTA keeps its chains' terms in shared memory and walks a contiguous range of octets; the octet rows
//     arrive through a 2-stage ring of bulk-TMA copies (full / empty mbarriers), so loading octet i+1 overlaps
//     computing octet i and no block barrier sits in the loop.
//   * the sums over events stay in registers (float64, fixed order) across the CTA's octets: one partial per
//     (chain, CTA) instead of one per (chain, 32 events), and one reduction per launch instead of per tile.
constexpr int kOct = 8;
constexpr int kQuad = 4;
struct OqSm {
  uint64_t* full;   // [2]
  uint64_t* empty;  // [2]
  float4* rows;     // [2][kOct][row]
  float4* cp;       // [nc][cps]
  float4* cpP;      // [nc][cps]
  float4* c0;       // [nc]  {-tc0, -ac0, -tc0', -ac0'}
  int row, cps;
};
__host__ __device__ inline int oq_cps(int S) { return (S / 2) | 1; }
__host__ __device__ inline size_t oq_smem(int S, int nc) {
  const int xrow = 2 + 4 * (S / 2);
  return 32 + (static_cast<size_t>(2) * kOct * (xrow + 1) + static_cast<size_t>(nc) * (2 * oq_cps(S) + 1)) * sizeof(float4);
}
__device__ __forceinline__ OqSm carve_oq_sm(unsigned char* base, int S, int xrow, int nc) {
  OqSm m;
  m.row = xrow + 1;
  m.cps = oq_cps(S);
  m.full = reinterpret_cast<uint64_t*>(base);
  m.empty = m.full + 2;
  m.rows = reinterpret_cast<float4*>(base + 32);
  m.cp = m.rows + 2 * kOct * m.row;
  m.cpP = m.cp + nc * m.cps;
  m.c0 = m.cpP + nc * m.cps;
  return m;
}

// the station terms of the CTA's nc chains (current and with the pending proposal applied), all threads
__device__ __forceinline__ void oq_stage_chain_terms(const OqSm& m, int nc, int c_base, int J, int S, const double* tc,
                                                     const double* ac, const int* which, const int* idx,
                                                     const double* xnew) {
  const int n_pairs = S / 2;
  for (int i = threadIdx.x; i < nc * n_pairs; i += blockDim.x) {
    const int lc = i / n_pairs, mm = i - lc * n_pairs, c = c_base + lc;
    if (c >= J) continue;
    const double* gtc = tc + static_cast<size_t>(c) * S;
    const double* gac = ac + static_cast<size_t>(c) * S;
    const int wh = which[c], pi = idx[c];
    const float pv = static_cast<float>(xnew[c]);
    const int j0 = 1 + 2 * mm, j1 = j0 + 1;
    float4 cur = make_float4(-static_cast<float>(gtc[j0]), 0.f, -static_cast<float>(gac[j0]), 0.f);
    if (j1 < S) {
      cur.y = -static_cast<float>(gtc[j1]);
      cur.w = -static_cast<float>(gac[j1]);
    }
    float4 prp = cur;
    if (wh == 2 && pi == j0) prp.x = -pv;
    if (wh == 2 && pi == j1) prp.y = -pv;
    if (wh == 4 && pi == j0) prp.z = -pv;
    if (wh == 4 && pi == j1) prp.w = -pv;
    m.cp[lc * m.cps + mm] = cur;
    m.cpP[lc * m.cps + mm] = prp;
  }
  for (int lc = threadIdx.x; lc < nc; lc += blockDim.x) {
    const int c = c_base + lc;
    if (c >= J) continue;
    const int wh = which[c], pi = idx[c];
    const float pv = static_cast<float>(xnew[c]);
    float4 c0 = make_float4(-static_cast<float>(tc[static_cast<size_t>(c) * S]), -static_cast<float>(ac[static_cast<size_t>(c) * S]),
                            0.f, 0.f);
    c0.z = (wh == 2 && pi == 0) ? -pv : c0.x;
    c0.w = (wh == 4 && pi == 0) ? -pv : c0.y;
    m.c0[lc] = c0;
  }
}

template <bool TRACE>
__global__ void __launch_bounds__(kCW * 32, 2) gibbs_sweep_oq_kernel(const GibbsParams<float> p, const GibbsDecide dec,
                                                                     unsigned int* done_counter, const int n_oct) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  const int S = p.S, J = p.J, E = p.E, n_pairs = S / 2;
  const int nc = n_warps * kQuad;
  const int c_base = blockIdx.y * nc;  // first chain of the CTA
  const OqSm m = carve_oq_sm(smem_raw, S, p.xrow, nc);
  const int o_begin = static_cast<int>(static_cast<long>(n_oct) * blockIdx.x / gridDim.x);
  const int o_end = static_cast<int>(static_cast<long>(n_oct) * (blockIdx.x + 1) / gridDim.x);
  const int n_my = o_end - o_begin;
  const uint32_t row_bytes = static_cast<uint32_t>(p.xrow * sizeof(float4));
  if (threadIdx.x == 0) {
    mbar_init(m.full, 1);
    mbar_init(m.full + 1, 1);
    mbar_init(m.empty, n_warps);
    mbar_init(m.empty + 1, n_warps);
    fence_mbar_init();
    fence_proxy_async();
  }
  __syncthreads();
  // one octet's rows -> ring stage `buf` (warp 0)
  auto issue = [&](int o, int buf) {
    const int n_ev = min(kOct, E - o * kOct);
    if (lane == 0) mbar_expect_tx(m.full + buf, row_bytes * n_ev);
    __syncwarp();
    if (lane < n_ev)
      tma_load_1d(m.rows + (buf * kOct + lane) * m.row, p.obsx + static_cast<size_t>(o * kOct + lane) * p.xrow, row_bytes,
                  m.full + buf);
  };
  if (warp == 0) {
    if (n_my > 0) issue(o_begin, 0);
    if (n_my > 1) issue(o_begin + 1, 1);
  }
  // the CTA's chain terms (once per launch)
  oq_stage_chain_terms(m, nc, c_base, J, S, p.g_tc, p.g_ac, p.prop_which, p.prop_idx, p.prop_xnew);
  __syncthreads();

  // this lane's chain (slots past chain J-1 clone the quad's first chain and never write)
  // the 4 chains of one event sit in adjacent lanes: a 16-byte shared-memory read is served per lane pair,
  // so lanes that share a row must be neighbours (measured: 2 wavefronts per row read against 4 for the
  // event-minor order; tools/micro/lds_pattern.cu)
  const int es = lane >> 2, chs = lane & (kQuad - 1);
  const bool warp_ok = c_base + warp * kQuad < J;
  const bool c_ok = c_base + warp * kQuad + chs < J;
  const int lc = c_ok ? warp * kQuad + chs : warp * kQuad;
  const int c = warp_ok ? c_base + lc : 0;
  StepIn<float> in;
  {
    const double Td = p.g_T[c];
    in.T = static_cast<float>(Td);
    in.iT = 1.f / in.T;
    in.cold = gibbs_is_cold<float>(Td);
    in.vs = static_cast<float>(p.g_vs[c]);
    in.qs = static_cast<float>(p.g_qs[c]);
    in.which = p.prop_which[c];
    in.pidx = p.prop_idx[c];
    in.pval = static_cast<float>(p.prop_xnew[c]);
    in.S = S;
    in.n_pairs = n_pairs;
    in.cp = m.cp + lc * m.cps;
    in.cpP = m.cpP + lc * m.cps;
    in.c0 = warp_ok ? m.c0[lc] : make_float4(0.f, 0.f, 0.f, 0.f);
    in.s_sta = nullptr;
    in.tc = nullptr;
    in.ac = nullptr;
  }
  const bool a_prev = p.a_prev[c] != 0;
  const int rec_chain_slot = (p.rec_slot >= 0 && p.hypo_rec) ? p.slot_of[c] : -1;
  double s_cur = 0.0, s_prop = 0.0;
  uint32_t cnt_p[3] = {0, 0, 0}, cnt_a[3] = {0, 0, 0};

  for (int i = 0; i < n_my; ++i) {
    const int buf = i & 1;
    const uint32_t ph = (i >> 1) & 1;
    const int o = o_begin + i;
    mbar_wait(m.full + buf, ph);
    if (warp_ok) {
      const int n_ev = min(kOct, E - o * kOct);
      const int e = o * kOct + es;
      const bool ev_ok = e < E;
      const int ee = ev_ok ? e : E - 1;  // idle lanes clone the last event and never write
      in.obs_row = m.rows + (buf * kOct + (ev_ok ? es : n_ev - 1)) * m.row;
      const size_t ci = static_cast<size_t>(c) * E + ee;
      const float4 evc = p.evc4[ee];
      const float mux = reinterpret_cast<const float*>(p.prior_xy)[2 * ee], muy = reinterpret_cast<const float*>(p.prior_xy)[2 * ee + 1];
      float x = p.hx[ci], y = p.hy[ci], z = p.hz[ci];
      float Le = a_prev ? p.hLp[ci] : p.hLe[ci];  // lazy commit of the last shared-parameter acceptance
      float Lp;
      int icmp;
      bool acc;
      gibbs_thread_step<float, TRACE>(p, p.it, in, c, e, ee, ev_ok && c_ok, evc, mux, muy, x, y, z, Le, Lp, icmp, acc, p.trace);
      if (ev_ok && c_ok) {
        p.hx[ci] = x;
        p.hy[ci] = y;
        p.hz[ci] = z;
        p.hLe[ci] = Le;
        p.hLp[ci] = Lp;
        if (rec_chain_slot >= 0)
          p.hypo_rec[(static_cast<size_t>(p.rec_slot) * p.n_cool_total + rec_chain_slot) * E + e] = make_float4(x, y, z, Le);
        s_cur += static_cast<double>(Le);
        s_prop += static_cast<double>(Lp);
        if (in.cold) {
#pragma unroll
          for (int t = 0; t < 3; ++t) {
            cnt_p[t] += icmp == t ? 1u : 0u;
            cnt_a[t] += (icmp == t && acc) ? 1u : 0u;
          }
        }
      }
    }
    // release the stage; warp 0 refills it with octet i+2 once every warp has let go of it
    __syncwarp();
    if (lane == 0) mbar_arrive(m.empty + buf);
    if (warp == 0 && i + 2 < n_my) {
      if (lane == 0) mbar_wait(m.empty + buf, ph);
      __syncwarp();
      issue(o + 2, buf);
    }
  }

  // per-(chain, CTA) partial sums: the 8 event lanes of a chain in a fixed order
  if (warp_ok) {
#pragma unroll
    for (int off = 16; off >= kQuad; off >>= 1) {
      s_cur += __shfl_xor_sync(0xffffffffu, s_cur, off);
      s_prop += __shfl_xor_sync(0xffffffffu, s_prop, off);
    }
    if (es == 0 && c_ok) {
      p.part_cur[static_cast<size_t>(c) * p.n_tiles + blockIdx.x] = s_cur;
      p.part_prop[static_cast<size_t>(c) * p.n_tiles + blockIdx.x] = s_prop;
    }
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const uint32_t np = warp_sum<uint32_t>(cnt_p[t]), na = warp_sum<uint32_t>(cnt_a[t]);
      if (lane == 0 && p.counts) {
        if (np) atomicAdd(p.counts + 4 + t, static_cast<unsigned long long>(np));
        if (na) atomicAdd(p.counts + 11 + t, static_cast<unsigned long long>(na));
      }
    }
  }

  last_cta_decides(dec, done_counter, smem_raw);
}

// ---- persistent cooperative kernel: all iterations in one launch -------------------------------------------
// Every (chain, event) keeps its state in registers, the tile's rows stay in shared memory, the chain-level
// state of ALL chains lives in every CTA's shared memory.  Per iteration: step -> partial sums to global ->
// ONE grid barrier -> every CTA adds the partials in the same fixed order and takes the same decisions
// (CTA (0,0) alone writes counters, records, traces).  Bit-identical to the per-iteration path.
template <typename real, bool TRACE>
__global__ void __launch_bounds__(kCW * 32, 2) gibbs_persist_kernel(const GibbsParams<real> p, const GibbsDecide d,
                                                                    const int iter_first, const int iter_last,
                                                                    const int rec_origin, const int rec_cap,
                                                                    htm_step_trace* trace_base, htm_swap_trace* swap_base,
                                                                    double* part /* [2][2][J][n_tiles] */) {
  typedef typename M<real>::real4 real4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cg::grid_group grid = cg::this_grid();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.S, J = p.J, E = p.E;
  const int tile = blockIdx.x, e = tile * kTile + lane;
  const int c = blockIdx.y * kCW + warp;
  const bool writer = blockIdx.x == 0 && blockIdx.y == 0;
  const SweepSm<real> m = carve_sweep_sm<real>(smem_raw, S, p.xrow);
  const ChainSm cs = carve_chain_sm(m.end, J, S);
  const int n_ev = min(kTile, E - tile * kTile);
  if (threadIdx.x == 0) {
    mbar_init(m.bar, 1);
    fence_mbar_init();
    fence_proxy_async();
  }
  __syncthreads();
  if (warp == 0) stage_tile<real>(p, m, tile, n_ev);
  chain_load(d, cs);
  __syncthreads();
  mbar_wait(m.bar, 0);

  const bool chain_ok = c < J;
  const bool ev_ok = e < E;
  const int ee = ev_ok ? e : E - 1;
  const int cc = chain_ok ? c : J - 1;
  const size_t ci = static_cast<size_t>(cc) * E + ee;
  const real4 evc = p.evc4[ee];
  const real mux = reinterpret_cast<const real*>(p.prior_xy)[2 * ee], muy = reinterpret_cast<const real*>(p.prior_xy)[2 * ee + 1];
  const real4* obs_row = m.obs + (ev_ok ? lane : n_ev - 1) * m.row;
  real x = p.hx[ci], y = p.hy[ci], z = p.hz[ci], Le = p.hLe[ci], Lp = p.hLp[ci];
  const size_t per_it = static_cast<size_t>(E + 1) * J, psz = static_cast<size_t>(J) * p.n_tiles;

  for (int it = iter_first; it <= iter_last; ++it) {
    const int buf = it & 1;
    double* part_cur = part + static_cast<size_t>(buf) * 2 * psz;
    double* part_prop = part_cur + psz;
    const bool rec = p.n_interval > 1 && (it % p.n_interval) == 1;
    int rec_slot = rec ? (it - 1) / p.n_interval - rec_origin : -1;
    if (rec_slot >= rec_cap) rec_slot = -1;
    if (chain_ok) {
      stage_chain_terms<real>(m, warp, S, cs.tc + static_cast<size_t>(c) * S, cs.ac + static_cast<size_t>(c) * S,
                              cs.which[c], cs.idx[c], cs.xnew[c]);
      __syncwarp();
      const StepIn<real> in = make_step_in<real>(m, warp, S, cs.T[c], cs.vs[c], cs.qs[c], cs.which[c], cs.idx[c],
                                                 cs.xnew[c], obs_row);
      Le = cs.aprev[c] ? Lp : Le;  // lazy commit of the last shared-parameter acceptance
      int icmp;
      bool acc;
      gibbs_thread_step<real, TRACE>(p, it, in, c, e, ee, ev_ok, evc, mux, muy, x, y, z, Le, Lp, icmp, acc,
                                     trace_base ? trace_base + static_cast<size_t>(it - iter_first) * per_it : nullptr);
      if (ev_ok && rec_slot >= 0 && cs.slot[c] >= 0 && p.hypo_rec) {
        real4 r4;
        r4.x = x;
        r4.y = y;
        r4.z = z;
        r4.w = Le;
        p.hypo_rec[(static_cast<size_t>(rec_slot) * p.n_cool_total + cs.slot[c]) * E + e] = r4;
      }
      tile_sums_and_counts<real>(p, part_cur, part_prop, c, tile, ev_ok, in.cold, Le, Lp, icmp, acc);
    }
    grid.sync();
    decide_core(d, cs, it, it + 1, part_cur, part_prop, rec_slot,
                trace_base ? trace_base + static_cast<size_t>(it - iter_first) * per_it + static_cast<size_t>(E) * J : nullptr,
                swap_base ? swap_base + (it - iter_first) : nullptr, writer);
  }
  if (chain_ok && ev_ok) {
    p.hx[ci] = x;
    p.hy[ci] = y;
    p.hz[ci] = z;
    p.hLe[ci] = Le;
    p.hLp[ci] = Lp;
  }
  if (writer) chain_store(d, cs);
}

// ---- chain set-up ---------------------------------------------------------------------------------
struct GibbsInitChain {
  double *g_vs, *g_qs, *g_tc, *g_ac, *g_T;
  int S, J, K, n_cool, ladder, solve_tc, solve_ac;
  double prior_vs, prior_qs, prior_tc, width_tc, prior_ac, width_ac, temp_high;
  uint64_t seed;
  uint32_t chain_offset;
};
__global__ void gibbs_init_chain_kernel(const GibbsInitChain a) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.J) return;
  const int k = c % a.K;
  a.g_vs[c] = a.prior_vs;  // start at the prior mean, src/hypo_tremor_mcmc.f90:175-185
  a.g_qs[c] = a.prior_qs;
  for (int j = 0; j < a.S; ++j) {
    const u32x4 w = philox4x32_10(a.seed, static_cast<uint32_t>(j), a.chain_offset + static_cast<uint32_t>(c), PHX_INIT, 1u);
    a.g_tc[static_cast<size_t>(c) * a.S + j] = a.solve_tc ? a.prior_tc + gauss64(w.v[0], w.v[1]) * a.width_tc : a.prior_tc;
    a.g_ac[static_cast<size_t>(c) * a.S + j] = a.solve_ac ? a.prior_ac + gauss64(w.v[2], w.v[3]) * a.width_ac : a.prior_ac;
  }
  double T = 1.0;
  if (k >= a.n_cool) {
    if (a.ladder == HTM_LADDER_GEOMETRIC) {
      T = ::exp(::log(a.temp_high) * static_cast<double>(k - a.n_cool + 1) / static_cast<double>(a.K - a.n_cool));
    } else {  // src/hypo_tremor_mcmc.f90:205-206
      const u32x4 w = philox4x32_10(a.seed, 0u, a.chain_offset + static_cast<uint32_t>(c), PHX_TEMP, 1u);
      T = ::exp((M<double>::u_co(w.v[0]) * (1.0 - kEps64) + kEps64) * ::log(a.temp_high));
    }
  }
  a.g_T[c] = T;
}

// one thread per (chain, event): generate_model for the hypocentre + its log-likelihood
template <typename real>
__global__ void gibbs_init_hypo_kernel(const GibbsParams<real> p, uint64_t seed) {
  typedef typename M<real>::real4 real4;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(p.J) * p.E) return;
  const int c = static_cast<int>(i / p.E), e = static_cast<int>(i % p.E);
  const uint32_t gid = (static_cast<uint32_t>(e) + p.event_offset) * p.J_total + p.chain_offset + static_cast<uint32_t>(c);
  const u32x4 a = philox4x32_10(seed, 0u, gid, PHX_INIT, 0u);
  const u32x4 b = philox4x32_10(seed, 1u, gid, PHX_INIT, 0u);
  const real mux = reinterpret_cast<const real*>(p.prior_xy)[2 * e], muy = reinterpret_cast<const real*>(p.prior_xy)[2 * e + 1];
  const real x = mux + M<real>::gauss(a.v[0], a.v[1]) * p.width_xy;
  const real y = muy + M<real>::gauss(a.v[2], a.v[3]) * p.width_xy;
  const real z = p.prior_z + M<real>::sqrt(static_cast<real>(-2) * M<real>::log(M<real>::u_oo(b.v[0]))) * p.width_z;
  const Glob<real> g = make_glob<real>(static_cast<real>(p.g_vs[c]), static_cast<real>(p.g_qs[c]));
  // station terms straight from global memory (one-time cost)
  real ct = 0, ca = 0, S1t = 0, S1a = 0, S2 = 0;
  for (int j = 0; j < p.S; ++j) {
    const real4 st = p.sta4[j];
    const real4 ob = p.obs4[static_cast<size_t>(e) * p.S + j];
    real rt, ra;
    station_resid(x, y, z, g, st, ob, static_cast<real>(p.g_tc[static_cast<size_t>(c) * p.S + j]),
                  static_cast<real>(p.g_ac[static_cast<size_t>(c) * p.S + j]), rt, ra);
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
  p.hx[i] = x;
  p.hy[i] = y;
  p.hz[i] = z;
  p.hLe[i] = finish_loglik<real>(S1t, S2, S1a, static_cast<real>(0), p.evc4[e]);
  p.hLp[i] = p.hLe[i];
}

// g_L[c] = sum_e L_e (fixed order), one warp per chain
template <typename real>
__global__ void gibbs_total_kernel(const real* hLe, int E, int J, double* g_L) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= J) return;
  double a = 0.0;
  for (int e = lane; e < E; e += 32) a += static_cast<double>(hLe[static_cast<size_t>(c) * E + e]);
  a = warp_sum<double>(a);
  if (lane == 0) g_L[c] = a;
}

// ---- persistent octet sweep: every iteration of a run in one cooperative launch, any E ----------------------
// gibbs_sweep_oq_kernel with the iteration loop inside: the grid is one wave of CTAs by construction, so the
// per-iteration launch, the re-staging of the chain terms from global memory and the single-CTA decide tail
// are replaced by ONE grid barrier per iteration and a decide step taken redundantly (and identically) by
// every CTA on its own shared-memory copy of the chain-level state, as in gibbs_persist_kernel.  Hypocentre
// state stays in global memory (L2): a (chain, event) is always visited by the same thread.  The TMA ring runs
// across iteration boundaries (rows never change), so the first octets of iteration i+1 arrive during the
// barrier and decide step of iteration i.
template <bool TRACE>
__global__ void __launch_bounds__(kCW * 32, 2) gibbs_persist_oq_kernel(const GibbsParams<float> p, const GibbsDecide d,
                                                                       const int iter_first, const int iter_last,
                                                                       const int rec_origin, const int rec_cap,
                                                                       htm_step_trace* trace_base, htm_swap_trace* swap_base,
                                                                       double* part /* [2][2][J][gridDim.x] */,
                                                                       const int n_oct, double* totals /* [2*J] */) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cg::grid_group grid = cg::this_grid();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  const int S = p.S, J = p.J, E = p.E, n_pairs = S / 2;
  const int nc = n_warps * kQuad;
  const int c_base = blockIdx.y * nc;
  const bool writer = blockIdx.x == 0 && blockIdx.y == 0;
  const OqSm m = carve_oq_sm(smem_raw, S, p.xrow, nc);
  const ChainSm cs = carve_chain_sm(smem_raw + oq_smem(S, nc), J, S);
  const int o_begin = static_cast<int>(static_cast<long>(n_oct) * blockIdx.x / gridDim.x);
  const int o_end = static_cast<int>(static_cast<long>(n_oct) * (blockIdx.x + 1) / gridDim.x);
  const int n_my = o_end - o_begin;  // >= 1: the launcher never starts more CTAs per chain group than octets
  const long n_run = static_cast<long>(n_my) * (iter_last - iter_first + 1);  // octet visits of this CTA
  const uint32_t row_bytes = static_cast<uint32_t>(p.xrow * sizeof(float4));
  if (threadIdx.x == 0) {
    mbar_init(m.full, 1);
    mbar_init(m.full + 1, 1);
    mbar_init(m.empty, n_warps);
    mbar_init(m.empty + 1, n_warps);
    fence_mbar_init();
    fence_proxy_async();
  }
  __syncthreads();
  auto issue = [&](long t) {  // visit t of this CTA -> ring stage t & 1
    const int buf = static_cast<int>(t & 1), o = o_begin + static_cast<int>(t % n_my);
    const int n_ev = min(kOct, E - o * kOct);
    if (lane == 0) mbar_expect_tx(m.full + buf, row_bytes * n_ev);
    __syncwarp();
    if (lane < n_ev)
      tma_load_1d(m.rows + (buf * kOct + lane) * m.row, p.obsx + static_cast<size_t>(o * kOct + lane) * p.xrow, row_bytes,
                  m.full + buf);
  };
  if (warp == 0) {
    if (n_run > 0) issue(0);
    if (n_run > 1) issue(1);
  }
  chain_load(d, cs);
  __syncthreads();

  const int es = lane >> 2, chs = lane & (kQuad - 1);  // chain minor: see gibbs_sweep_oq_kernel
  const bool warp_ok = c_base + warp * kQuad < J;
  const bool c_ok = c_base + warp * kQuad + chs < J;
  const int lc = c_ok ? warp * kQuad + chs : warp * kQuad;
  const int c = warp_ok ? c_base + lc : 0;
  const size_t per_it = static_cast<size_t>(E + 1) * J, psz = static_cast<size_t>(J) * gridDim.x;
  uint32_t cnt_p[3] = {0, 0, 0}, cnt_a[3] = {0, 0, 0};
  long t_run = 0;

  for (int it = iter_first; it <= iter_last; ++it) {
    double* part_cur = part + static_cast<size_t>(it & 1) * 2 * psz;
    double* part_prop = part_cur + psz;
    const bool rec = p.n_interval > 1 && (it % p.n_interval) == 1;
    int rec_slot = rec ? (it - 1) / p.n_interval - rec_origin : -1;
    if (rec_slot >= rec_cap) rec_slot = -1;
    htm_step_trace* trace_it = trace_base ? trace_base + static_cast<size_t>(it - iter_first) * per_it : nullptr;
    // the CTA's chain terms for this iteration, from the shared-memory chain state (decide_core ended with a
    // block barrier; the previous iteration's reads of cp/cpP are over)
    oq_stage_chain_terms(m, nc, c_base, J, S, cs.tc, cs.ac, cs.which, cs.idx, cs.xnew);
    __syncthreads();
    StepIn<float> in;
    {
      const double Td = cs.T[c];
      in.T = static_cast<float>(Td);
      in.iT = 1.f / in.T;
      in.cold = gibbs_is_cold<float>(Td);
      in.vs = static_cast<float>(cs.vs[c]);
      in.qs = static_cast<float>(cs.qs[c]);
      in.which = cs.which[c];
      in.pidx = cs.idx[c];
      in.pval = static_cast<float>(cs.xnew[c]);
      in.S = S;
      in.n_pairs = n_pairs;
      in.cp = m.cp + lc * m.cps;
      in.cpP = m.cpP + lc * m.cps;
      in.c0 = warp_ok ? m.c0[lc] : make_float4(0.f, 0.f, 0.f, 0.f);
      in.s_sta = nullptr;
      in.tc = nullptr;
      in.ac = nullptr;
    }
    const bool a_prev = cs.aprev[c] != 0;
    const int rec_chain_slot = (rec_slot >= 0 && p.hypo_rec) ? cs.slot[c] : -1;
    double s_cur = 0.0, s_prop = 0.0;
    for (int i = 0; i < n_my; ++i, ++t_run) {
      const int buf = static_cast<int>(t_run & 1);
      const uint32_t ph = static_cast<uint32_t>((t_run >> 1) & 1);
      const int o = o_begin + i;
      mbar_wait(m.full + buf, ph);
      if (warp_ok) {
        const int n_ev = min(kOct, E - o * kOct);
        const int e = o * kOct + es;
        const bool ev_ok = e < E;
        const int ee = ev_ok ? e : E - 1;
        in.obs_row = m.rows + (buf * kOct + (ev_ok ? es : n_ev - 1)) * m.row;
        const size_t ci = static_cast<size_t>(c) * E + ee;
        const float4 evc = p.evc4[ee];
        const float mux = reinterpret_cast<const float*>(p.prior_xy)[2 * ee], muy = reinterpret_cast<const float*>(p.prior_xy)[2 * ee + 1];
        float x = p.hx[ci], y = p.hy[ci], z = p.hz[ci];
        float Le = a_prev ? p.hLp[ci] : p.hLe[ci];  // lazy commit of the last shared-parameter acceptance
        float Lp;
        int icmp;
        bool acc;
        gibbs_thread_step<float, TRACE>(p, it, in, c, e, ee, ev_ok && c_ok, evc, mux, muy, x, y, z, Le, Lp, icmp, acc, trace_it);
        if (ev_ok && c_ok) {
          p.hx[ci] = x;
          p.hy[ci] = y;
          p.hz[ci] = z;
          p.hLe[ci] = Le;
          p.hLp[ci] = Lp;
          if (rec_chain_slot >= 0)
            p.hypo_rec[(static_cast<size_t>(rec_slot) * p.n_cool_total + rec_chain_slot) * E + e] = make_float4(x, y, z, Le);
          s_cur += static_cast<double>(Le);
          s_prop += static_cast<double>(Lp);
          if (in.cold) {
#pragma unroll
            for (int t = 0; t < 3; ++t) {
              cnt_p[t] += icmp == t ? 1u : 0u;
              cnt_a[t] += (icmp == t && acc) ? 1u : 0u;
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(m.empty + buf);
      if (warp == 0 && t_run + 2 < n_run) {
        if (lane == 0) mbar_wait(m.empty + buf, ph);
        __syncwarp();
        issue(t_run + 2);
      }
    }
    if (warp_ok) {
#pragma unroll
      for (int off = 16; off >= kQuad; off >>= 1) {
        s_cur += __shfl_xor_sync(0xffffffffu, s_cur, off);
        s_prop += __shfl_xor_sync(0xffffffffu, s_prop, off);
      }
      if (es == 0 && c_ok) {
        part_cur[static_cast<size_t>(c) * gridDim.x + blockIdx.x] = s_cur;
        part_prop[static_cast<size_t>(c) * gridDim.x + blockIdx.x] = s_prop;
      }
    }
    grid.sync();
    htm_step_trace* trace_g = trace_it ? trace_it + static_cast<size_t>(E) * J : nullptr;
    htm_swap_trace* swap_it = swap_base ? swap_base + (it - iter_first) : nullptr;
    if (d.xch.n > 1) {
      // event shards: CTA (0,0) adds this shard's partials, exchanges the sums with the other GPUs through peer
      // memory and hands the totals over all events to every CTA of its grid
      if (writer) {
        sum_partials(d.n_tiles, part_cur, part_prop, J, cs.tot);
        __syncthreads();
        if (peer_allreduce(d.xch, d.xch.epoch + static_cast<uint32_t>(it - iter_first), cs.tot, 2 * J))
          for (int t = threadIdx.x; t < 2 * J; t += blockDim.x) totals[t] = cs.tot[t];
        __threadfence();
      }
      grid.sync();
      // a peer that never answered ends the run here, on every CTA alike (only this shard's writer sets the
      // flag, before the barrier): no decision is taken from partial sums; the host reports HTM_ERR_CUDA
      if (*reinterpret_cast<volatile int*>(d.xch.status) != 0) break;
      decide_core(d, cs, it, it + 1, totals, totals + J, rec_slot, trace_g, swap_it, writer, true);
    } else {
      decide_core(d, cs, it, it + 1, part_cur, part_prop, rec_slot, trace_g, swap_it, writer);
    }
  }
  if (warp_ok) {
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const uint32_t np = warp_sum<uint32_t>(cnt_p[t]), na = warp_sum<uint32_t>(cnt_a[t]);
      if (lane == 0 && p.counts) {
        if (np) atomicAdd(p.counts + 4 + t, static_cast<unsigned long long>(np));
        if (na) atomicAdd(p.counts + 11 + t, static_cast<unsigned long long>(na));
      }
    }
  }
  if (writer) chain_store(d, cs);
}

// ---- host launchers ---------------------------------------------------------------------------------
template <typename real>
static GibbsParams<real> make_gibbs_params(const GibbsLaunch& a) {
  GibbsParams<real> p;
  typedef typename M<real>::real4 real4;
  p.sta4 = static_cast<const real4*>(a.tab.sta4);
  p.obs4 = static_cast<const real4*>(a.tab.obs4_raw);
  p.evc4 = static_cast<const real4*>(a.tab.evc4);
  p.prior_xy = a.tab.prior_xy;
  p.obsx = static_cast<const float4*>(a.obsx);
  p.xrow = a.xrow;
  p.hx = static_cast<real*>(a.hx);
  p.hy = static_cast<real*>(a.hy);
  p.hz = static_cast<real*>(a.hz);
  p.hLe = static_cast<real*>(a.hLe);
  p.hLp = static_cast<real*>(a.hLp);
  p.g_vs = a.g_vs;
  p.g_qs = a.g_qs;
  p.g_tc = a.g_tc;
  p.g_ac = a.g_ac;
  p.g_T = a.g_T;
  p.g_L = a.g_L;
  p.prop_which = a.prop_which;
  p.prop_idx = a.prop_idx;
  p.prop_xnew = a.prop_xnew;
  p.prop_lpr = a.prop_lpr;
  p.a_prev = a.a_prev;
  p.slot_of = a.slot_of;
  p.part_cur = a.part_cur;  // one allocation [2][2][J][n_tiles]; the per-iteration path uses buffer 0
  p.part_prop = a.part_cur + static_cast<size_t>(a.J) * ((a.E + kTile - 1) / kTile);
  p.E = a.E;
  p.S = a.S;
  p.J = a.J;
  p.K = a.K;
  p.n_tiles = (a.E + kTile - 1) / kTile;
  p.n_cool_total = a.n_cool_total;
  p.it = 0;
  p.n_burn = a.n_burn;
  p.n_interval = a.n_interval;
  p.rk = philox_keys(a.seed);
  p.event_offset = a.event_offset;
  p.chain_offset = a.chain_offset;
  p.J_total = a.J_total;
  p.prior_z = static_cast<real>(a.prior_z);
  p.width_z = static_cast<real>(a.width_z);
  p.width_xy = static_cast<real>(a.width_xy);
  p.step_xy = static_cast<real>(a.step_xy);
  p.step_z = static_cast<real>(a.step_z);
  p.counts = a.counts;
  p.hypo_rec = static_cast<real4*>(a.hypo_rec);
  p.rec_slot = -1;
  p.trace = nullptr;
  p.swap = nullptr;
  return p;
}

static GibbsDecide make_decide(const GibbsLaunch& a) {
  GibbsDecide d;
  d.g_vs = a.g_vs;
  d.g_qs = a.g_qs;
  d.g_tc = a.g_tc;
  d.g_ac = a.g_ac;
  d.g_T = a.g_T;
  d.g_L = a.g_L;
  d.prop_which = a.prop_which;
  d.prop_idx = a.prop_idx;
  d.prop_xnew = a.prop_xnew;
  d.prop_lpr = a.prop_lpr;
  d.a_prev = a.a_prev;
  d.slot_of = a.slot_of;
  d.part_cur = a.part_cur;
  d.part_prop = a.part_cur + static_cast<size_t>(a.J) * ((a.E + kTile - 1) / kTile);
  d.S = a.S;
  d.J = a.J;
  d.K = a.K;
  d.n_tiles = (a.E + kTile - 1) / kTile;
  d.n_cool_total = a.n_cool_total;
  d.it = 0;
  d.it_next = 0;
  d.rk = philox_keys(a.seed);
  d.chain_offset = a.chain_offset;
  d.swap_stream = a.swap_stream;
  d.n_solved = 0;
  for (int t = 0; t < 4; ++t) {
    d.solved[t] = 0;
    if (a.solve[t]) d.solved[d.n_solved++] = t + 1;
    d.prior[t] = a.g_prior[t];
    d.width[t] = a.g_width[t];
    d.step[t] = a.g_step[t];
  }
  d.counts = a.counts;
  d.count_globals = a.count_globals;
  d.xch = a.xch;
  d.rec_slot = -1;
  d.rec_chain = a.rec_chain;
  d.rec_vs = a.rec_vs;
  d.rec_qs = a.rec_qs;
  d.rec_L = a.rec_L;
  d.rec_tc = a.rec_tc;
  d.rec_ac = a.rec_ac;
  d.trace = nullptr;
  d.swap = nullptr;
  return d;
}

// per-chain sums of the per-tile partials of THIS shard (fixed order): totals[c] = cur, totals[J + c] = proposed
__global__ void gibbs_totals_kernel(const double* part_cur, const double* part_prop, int J, int n_tiles, double* totals) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= J) return;
  double a = 0.0, b = 0.0;
  for (int t = lane; t < n_tiles; t += 32) {
    a += part_cur[static_cast<size_t>(c) * n_tiles + t];
    b += part_prop[static_cast<size_t>(c) * n_tiles + t];
  }
  a = warp_sum<double>(a);
  b = warp_sum<double>(b);
  if (lane == 0) {
    totals[c] = a;
    totals[J + c] = b;
  }
}

// HTM_GIBBS_PERSIST=0 forces one launch per iteration, =1 insists on the persistent kernel (tests)
static int persist_env() {
  const char* v = std::getenv("HTM_GIBBS_PERSIST");
  return v ? std::atoi(v) : -1;
}
// HTM_GIBBS_SWEEP=chain|octet fixes the layout of the float32 sweeps, per-iteration and persistent (tests, tuning)
static int sweep_env() {
  const char* v = std::getenv("HTM_GIBBS_SWEEP");
  if (!v) return -1;
  return std::string(v) == "octet" ? 1 : 0;
}

// One sweep launch of the per-iteration paths
struct SweepShape {
  bool octet = false;  // gibbs_sweep_oq_kernel (float32) instead of gibbs_sweep_kernel
  int n_warps = kCW;
  dim3 grid;
  size_t smem = 0;  // sweep part only
};
template <typename real, bool TRACE>
static cudaError_t sweep_set_smem(const SweepShape& s, size_t smem) {
  const void* f = reinterpret_cast<const void*>(gibbs_sweep_kernel<real, TRACE>);
  if (s.octet) f = reinterpret_cast<const void*>(gibbs_sweep_oq_kernel<TRACE>);
  return cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
}
template <typename real, bool TRACE>
static void sweep_launch(const SweepShape& s, size_t smem, cudaStream_t stream, const GibbsParams<real>& p,
                         const GibbsDecide& d, unsigned int* done_counter) {
  if constexpr (sizeof(real) == 4) {
    if (s.octet) {
      gibbs_sweep_oq_kernel<TRACE><<<s.grid, s.n_warps * 32, smem, stream>>>(p, d, done_counter, (p.E + kOct - 1) / kOct);
      return;
    }
  }
  gibbs_sweep_kernel<real, TRACE><<<s.grid, kCW * 32, smem, stream>>>(p, d, done_counter);
}

template <typename real, bool TRACE>
static cudaError_t launch_gibbs_tt(const GibbsLaunch& a, cudaStream_t stream, int* n_launches) {
  GibbsParams<real> p = make_gibbs_params<real>(a);
  GibbsDecide d = make_decide(a);
  const size_t sm_sweep = sweep_smem<real>(a.S), sm_chain = chain_sm_bytes(a.J, a.S);
  if (sm_sweep > 200 * 1024 || sm_chain > 200 * 1024) return cudaErrorInvalidConfiguration;
  SweepShape shape;
  shape.grid = dim3(p.n_tiles, (a.J + kCW - 1) / kCW);
  shape.smem = sm_sweep;
  // the last CTA reuses its shared memory for the decide step
  size_t smem_iter = sm_sweep > sm_chain ? sm_sweep : sm_chain;
  GibbsParams<real> ps = p;  // what the per-iteration sweeps see
  GibbsDecide ds = d;
  cudaError_t err = cudaSuccess;
  if constexpr (sizeof(real) == 4) {
    // octet layout: one wave of CTAs, each walking a contiguous range of event octets with the terms of up to
    // 32 chains in shared memory; worth its set-up cost once every CTA gets a few octets
    const int quads = (a.J + kQuad - 1) / kQuad;
    const int gy = (quads + kCW - 1) / kCW;
    const int n_warps = (quads + gy - 1) / gy;
    const size_t sm_oq = oq_smem(a.S, n_warps * kQuad);
    const size_t smem_oq = sm_oq > sm_chain ? sm_oq : sm_chain;
    const int want = sweep_env();
    if (want != 0 && smem_oq <= 200 * 1024) {
      err = cudaFuncSetAttribute(gibbs_sweep_oq_kernel<TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem_oq));
      if (err != cudaSuccess) return err;
      int per_sm = 0, dev = 0, n_sm = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
      err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gibbs_sweep_oq_kernel<TRACE>, n_warps * 32, smem_oq);
      if (err != cudaSuccess) return err;
      const long n_oct = (a.E + kOct - 1) / kOct;
      long gx = static_cast<long>(per_sm) * n_sm / gy;
      const bool pays = gx >= 1 && n_oct >= 3 * gx;
      if (gx > n_oct) gx = n_oct;
      if (gx > 2L * a.part_tiles) gx = 2L * a.part_tiles;  // [cur, prop][J][gx] must fit the 4 * J * part_tiles buffer
      if (gx >= 1 && (want == 1 || pays)) {
        shape.octet = true;
        shape.n_warps = n_warps;
        shape.grid = dim3(static_cast<unsigned>(gx), gy);
        shape.smem = sm_oq;
        smem_iter = smem_oq;
        ps.n_tiles = ds.n_tiles = static_cast<int>(gx);  // one partial sum per (chain, CTA)
        ps.part_prop = ps.part_cur + static_cast<size_t>(a.J) * gx;
        ds.part_prop = ps.part_prop;
      }
    }
  }
  const size_t smem_pers = sm_sweep + sm_chain;
  const dim3 grid(p.n_tiles, (a.J + kCW - 1) / kCW);  // persistent kernel: one chain per warp
  err = cudaFuncSetAttribute(gibbs_decide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm_chain));
  if (err != cudaSuccess) return err;
  int nl = 0;
  // prepare: cold slots + the proposal of the first iteration (a pure function of state and iteration)
  d.it = 0;
  d.it_next = a.iter_first;
  gibbs_decide_kernel<<<1, 256, sm_chain, stream>>>(d);
  ++nl;

  const bool peer_xch = a.xch.n > 1;  // one launch per iteration; its last CTA exchanges the sums and decides
  // ---- event-sharded joint chains: sweep -> local totals -> all-reduce -> decide (replicated) ----
  if (a.comm && !peer_xch) {
    err = sweep_set_smem<real, TRACE>(shape, smem_iter);
    if (err != cudaSuccess) return err;
    const size_t per_it_s = static_cast<size_t>(a.E + 1) * a.J;
    std::string why;
    for (int it = a.iter_first; it <= a.iter_last; ++it) {
      const bool rec = a.n_interval > 1 && (it % a.n_interval) == 1;
      const int slot = rec ? (it - 1) / a.n_interval - a.rec_origin : -1;
      ps.it = it;
      ps.rec_slot = (slot >= 0 && slot < a.rec_cap) ? slot : -1;
      ps.trace = a.trace ? a.trace + static_cast<size_t>(it - a.iter_first) * per_it_s : nullptr;
      sweep_launch<real, TRACE>(shape, smem_iter, stream, ps, ds, nullptr);
      gibbs_totals_kernel<<<(a.J + 3) / 4, 128, 0, stream>>>(ps.part_cur, ps.part_prop, a.J, ps.n_tiles, a.totals);
      if (!nccl_allreduce_f64(a.comm, a.totals, a.totals, 2 * static_cast<size_t>(a.J), stream, &why))
        return cudaErrorUnknown;
      GibbsDecide dd = ds;
      dd.n_tiles = 1;  // the "partials" are now the global per-chain sums
      dd.part_cur = a.totals;
      dd.part_prop = a.totals + a.J;
      dd.it = it;
      dd.it_next = it + 1;
      dd.rec_slot = ps.rec_slot;
      dd.trace = a.trace ? a.trace + static_cast<size_t>(it - a.iter_first) * per_it_s + static_cast<size_t>(a.E) * a.J : nullptr;
      dd.swap = a.swaps ? a.swaps + (it - a.iter_first) : nullptr;
      gibbs_decide_kernel<<<1, 256, sm_chain, stream>>>(dd);
      nl += 3;
    }
    if (n_launches) *n_launches = nl;
    return cudaGetLastError();
  }

  // ---- persistent cooperative kernel when every CTA can be resident at once ----
  const int want = persist_env();
  bool persistent = false;
  if (want != 0 && !peer_xch && smem_pers <= 200 * 1024 && !(sizeof(real) == 4 && sweep_env() == 1)) {  // HTM_GIBBS_SWEEP=octet skips it
    err = cudaFuncSetAttribute(gibbs_persist_kernel<real, TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem_pers));
    if (err != cudaSuccess) return err;
    int per_sm = 0, dev = 0, n_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gibbs_persist_kernel<real, TRACE>, kCW * 32, smem_pers);
    if (err != cudaSuccess) return err;
    persistent = static_cast<long>(per_sm) * n_sm >= static_cast<long>(grid.x) * grid.y;
  }
  // float32, too many tiles for that: the persistent octet sweep (one wave of CTAs walking event octets)
  if constexpr (sizeof(real) == 4) {
    if (!persistent && want != 0 && sweep_env() != 0) {  // also the event-sharded run with peer-memory exchange
      const int quads = (a.J + kQuad - 1) / kQuad;
      const int gy = (quads + kCW - 1) / kCW;
      const int n_warps = (quads + gy - 1) / gy;
      const size_t smem_po = oq_smem(a.S, n_warps * kQuad) + sm_chain;
      if (smem_po <= 200 * 1024) {
        err = cudaFuncSetAttribute(gibbs_persist_oq_kernel<TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem_po));
        if (err != cudaSuccess) return err;
        int per_sm = 0, dev = 0, n_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gibbs_persist_oq_kernel<TRACE>, n_warps * 32, smem_po);
        if (err != cudaSuccess) return err;
        int n_oct = (a.E + kOct - 1) / kOct;
        long gx = static_cast<long>(per_sm) * n_sm / gy;
        if (gx > n_oct) gx = n_oct;
        if (gx > a.part_tiles) gx = a.part_tiles;  // [2][cur, prop][J][gx]
        if (gx >= 1) {
          GibbsParams<real> pp = p;
          GibbsDecide dp = d;
          pp.n_tiles = dp.n_tiles = static_cast<int>(gx);
          int iter_first = a.iter_first, iter_last = a.iter_last, rec_origin = a.rec_origin, rec_cap = a.rec_cap;
          htm_step_trace* tr = a.trace;
          htm_swap_trace* sw = a.swaps;
          double* part = a.part_cur;
          double* totals = a.totals;
          dp.xch.epoch = a.xch_epoch0;  // exchange number of iter_first; the kernel counts on from there
          void* args[] = {&pp, &dp, &iter_first, &iter_last, &rec_origin, &rec_cap, &tr, &sw, &part, &n_oct, &totals};
          err = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(gibbs_persist_oq_kernel<TRACE>),
                                            dim3(static_cast<unsigned>(gx), gy), dim3(n_warps * 32), args, smem_po, stream);
          if (err != cudaSuccess) return err;
          ++nl;
          if (n_launches) *n_launches = nl;
          return cudaGetLastError();
        }
      }
    }
  }
  if (want == 1 && !persistent && !peer_xch) return cudaErrorCooperativeLaunchTooLarge;
  if (persistent) {
    int iter_first = a.iter_first, iter_last = a.iter_last, rec_origin = a.rec_origin, rec_cap = a.rec_cap;
    htm_step_trace* tr = a.trace;
    htm_swap_trace* sw = a.swaps;
    double* part = a.part_cur;  // [2][2][J][n_tiles]: part_cur and part_prop are one allocation
    void* args[] = {&p, &d, &iter_first, &iter_last, &rec_origin, &rec_cap, &tr, &sw, &part};
    err = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(gibbs_persist_kernel<real, TRACE>), grid, dim3(kCW * 32),
                                      args, smem_pers, stream);
    if (err != cudaSuccess) return err;
    ++nl;
    if (n_launches) *n_launches = nl;
    return cudaGetLastError();
  }

  // ---- one launch per iteration: the sweep, and in its last CTA the chain-level decide step ----
  err = sweep_set_smem<real, TRACE>(shape, smem_iter);
  if (err != cudaSuccess) return err;
  const size_t per_it = static_cast<size_t>(a.E + 1) * a.J;
  for (int it = a.iter_first; it <= a.iter_last; ++it) {
    const bool rec = a.n_interval > 1 && (it % a.n_interval) == 1;
    const int slot = rec ? (it - 1) / a.n_interval - a.rec_origin : -1;
    ps.it = it;
    ps.rec_slot = (slot >= 0 && slot < a.rec_cap) ? slot : -1;
    ps.trace = a.trace ? a.trace + static_cast<size_t>(it - a.iter_first) * per_it : nullptr;
    ds.it = it;
    ds.it_next = it + 1;
    ds.rec_slot = ps.rec_slot;
    ds.trace = a.trace ? a.trace + static_cast<size_t>(it - a.iter_first) * per_it + static_cast<size_t>(a.E) * a.J : nullptr;
    ds.swap = a.swaps ? a.swaps + (it - a.iter_first) : nullptr;
    ds.xch.epoch = a.xch_epoch0 + static_cast<uint32_t>(it - a.iter_first);
    sweep_launch<real, TRACE>(shape, smem_iter, stream, ps, ds, a.done_counter);
    ++nl;
  }
  if (n_launches) *n_launches = nl;
  return cudaGetLastError();
}

template <typename real>
static cudaError_t launch_gibbs_t(const GibbsLaunch& a, cudaStream_t stream, int* n_launches) {
  return (a.trace || a.swaps) ? launch_gibbs_tt<real, true>(a, stream, n_launches)
                              : launch_gibbs_tt<real, false>(a, stream, n_launches);
}

cudaError_t launch_gibbs(const GibbsLaunch& a, cudaStream_t stream, int* n_launches) {
  return a.precision == HTM_PRECISION_F64 ? launch_gibbs_t<double>(a, stream, n_launches)
                                          : launch_gibbs_t<float>(a, stream, n_launches);
}

cudaError_t launch_gibbs_init(const GibbsLaunch& a, double temp_high, int ladder, int n_cool, cudaStream_t stream) {
  GibbsInitChain ic;
  ic.g_vs = a.g_vs;
  ic.g_qs = a.g_qs;
  ic.g_tc = a.g_tc;
  ic.g_ac = a.g_ac;
  ic.g_T = a.g_T;
  ic.S = a.S;
  ic.J = a.J;
  ic.K = a.K;
  ic.n_cool = n_cool;
  ic.ladder = ladder;
  ic.solve_tc = a.solve[1];
  ic.solve_ac = a.solve[3];
  ic.prior_vs = a.g_prior[0];
  ic.prior_tc = a.g_prior[1];
  ic.prior_qs = a.g_prior[2];
  ic.prior_ac = a.g_prior[3];
  ic.width_tc = a.g_width[1];
  ic.width_ac = a.g_width[3];
  ic.temp_high = temp_high;
  ic.seed = a.seed;
  ic.chain_offset = a.chain_offset;
  gibbs_init_chain_kernel<<<(a.J + 63) / 64, 64, 0, stream>>>(ic);
  const size_t n = static_cast<size_t>(a.J) * a.E;
  const unsigned grid = static_cast<unsigned>((n + 127) / 128);
  if (a.precision == HTM_PRECISION_F64) {
    gibbs_init_hypo_kernel<double><<<grid, 128, 0, stream>>>(make_gibbs_params<double>(a), a.seed);
    gibbs_total_kernel<double><<<(a.J + 3) / 4, 128, 0, stream>>>(static_cast<const double*>(a.hLe), a.E, a.J, a.g_L);
  } else {
    gibbs_init_hypo_kernel<float><<<grid, 128, 0, stream>>>(make_gibbs_params<float>(a), a.seed);
    gibbs_total_kernel<float><<<(a.J + 3) / 4, 128, 0, stream>>>(static_cast<const float*>(a.hLe), a.E, a.J, a.g_L);
  }
  return cudaGetLastError();
}

}  // namespace htm
