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
// This file holds the FLOAT64 kernels: the validation path, which keeps the reference's operation order station
// by station and is compared step by step with the oracle.  The float32 throughput kernel (same schedule,
// moment-based evaluation of the shared-parameter proposals) is htm_gibbs_f32.cu.
//
// Mapping of the sweep: CTA = tile of 32 events x up to 8 chains; warp = chain, lane = event.  The
// tile's observation rows are staged by 32 bulk-TMA copies (one per event) into padded shared-memory
// rows, so the per-lane 16-byte reads are bank-conflict free; station table and the chains' station
// terms are warp-broadcast reads.  The persistent cooperative kernel takes over whenever every tile can be
// resident at once.
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


// shared-memory carve-up common to both kernels
template <typename real>
struct SweepSm {
  typedef typename M<real>::real4 real4;
  uint64_t* bar;
  real4* obs;  // [kTile][row]
  real4* sta;  // [S]
  real* tc;    // [kCW][S]
  real* ac;    // [kCW][S]
  int row;
  unsigned char* end;  // first byte after the sweep's own shared memory (16-byte aligned)
};
template <typename real>
__device__ __forceinline__ SweepSm<real> carve_sweep_sm(unsigned char* base, int S) {
  typedef typename M<real>::real4 real4;
  SweepSm<real> m;
  m.row = S + 1;  // padded: the per-lane 16/32-byte row reads are bank-conflict free
  m.bar = reinterpret_cast<uint64_t*>(base);
  m.obs = reinterpret_cast<real4*>(base + 16);
  m.sta = m.obs + kTile * m.row;
  m.tc = reinterpret_cast<real*>(m.sta + S);
  m.ac = m.tc + kCW * S;
  unsigned char* e = reinterpret_cast<unsigned char*>(m.ac + kCW * S);
  m.end = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(e) + 15) & ~static_cast<uintptr_t>(15));
  return m;
}
template <typename real>
static size_t sweep_smem(int S) {
  typedef typename M<real>::real4 real4;
  return 16 + static_cast<size_t>(kTile) * (S + 1) * sizeof(real4) + S * sizeof(real4) + 2 * kCW * S * sizeof(real) + 16;
}

// TMA-stage the tile's observation rows and the station table; caller syncs and waits
template <typename real>
__device__ __forceinline__ void stage_tile(const GibbsParams<real>& p, const SweepSm<real>& m, int tile, int n_ev) {
  typedef typename M<real>::real4 real4;
  const int lane = threadIdx.x & 31;
  const uint32_t bytes = static_cast<uint32_t>(p.S * sizeof(real4));
  if (lane == 0) mbar_expect_tx(m.bar, bytes * (n_ev + 1));
  __syncwarp();
  if (lane < n_ev)
    tma_load_1d(m.obs + lane * m.row, p.obs4 + static_cast<size_t>(tile * kTile + lane) * p.S, bytes, m.bar);
  if (lane == 0) tma_load_1d(m.sta, p.sta4, bytes, m.bar);
}

// the warp's chain operands in shared memory, from (tc, ac) arrays of any addressable memory
template <typename real>
__device__ __forceinline__ void stage_chain_terms(const SweepSm<real>& m, int warp, int S, const double* gtc,
                                                  const double* gac) {
  const int lane = threadIdx.x & 31;
  for (int j = lane; j < S; j += 32) {
    m.tc[warp * S + j] = static_cast<real>(gtc[j]);
    m.ac[warp * S + j] = static_cast<real>(gac[j]);
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
  in.obs_row = obs_row;
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
  const SweepSm<real> m = carve_sweep_sm<real>(smem_raw, S);
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
    stage_chain_terms<real>(m, warp, S, p.g_tc + static_cast<size_t>(c) * S, p.g_ac + static_cast<size_t>(c) * S);
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
  const SweepSm<real> m = carve_sweep_sm<real>(smem_raw, S);
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
      stage_chain_terms<real>(m, warp, S, cs.tc + static_cast<size_t>(c) * S, cs.ac + static_cast<size_t>(c) * S);
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

// ---- host launchers (make_gibbs_params / make_decide: htm_gibbs_decide.cuh) -------------------------------
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

// cold slots + the proposal of iteration a.iter_first (a pure function of state and iteration number)
cudaError_t launch_gibbs_prepare_f64(const GibbsLaunch& a, cudaStream_t stream) {
  const size_t sm_chain = chain_sm_bytes(a.J, a.S);
  cudaError_t err = cudaFuncSetAttribute(gibbs_decide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm_chain));
  if (err != cudaSuccess) return err;
  GibbsDecide d = make_decide(a);
  d.it = 0;
  d.it_next = a.iter_first;
  gibbs_decide_kernel<<<1, 256, sm_chain, stream>>>(d);
  return cudaGetLastError();
}

template <typename real, bool TRACE>
static cudaError_t launch_gibbs_tt(const GibbsLaunch& a, cudaStream_t stream, int* n_launches) {
  GibbsParams<real> p = make_gibbs_params<real>(a);
  GibbsDecide d = make_decide(a);
  const size_t sm_sweep = sweep_smem<real>(a.S), sm_chain = chain_sm_bytes(a.J, a.S);
  if (sm_sweep > 200 * 1024 || sm_chain > 200 * 1024) return cudaErrorInvalidConfiguration;
  const dim3 grid(p.n_tiles, (a.J + kCW - 1) / kCW);  // one chain per warp, one 32-event tile per CTA
  // the last CTA reuses its shared memory for the decide step
  const size_t smem_iter = sm_sweep > sm_chain ? sm_sweep : sm_chain;
  const size_t smem_pers = sm_sweep + sm_chain;
  cudaError_t err = launch_gibbs_prepare_f64(a, stream);
  if (err != cudaSuccess) return err;
  int nl = 1;

  const bool peer_xch = a.xch.n > 1;  // one launch per iteration; its last CTA exchanges the sums and decides
  // ---- event-sharded joint chains: sweep -> local totals -> all-reduce -> decide (replicated) ----
  if (a.comm && !peer_xch) {
    err = cudaFuncSetAttribute(gibbs_sweep_kernel<real, TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem_iter));
    if (err != cudaSuccess) return err;
    const size_t per_it_s = static_cast<size_t>(a.E + 1) * a.J;
    std::string why;
    for (int it = a.iter_first; it <= a.iter_last; ++it) {
      const bool rec = a.n_interval > 1 && (it % a.n_interval) == 1;
      const int slot = rec ? (it - 1) / a.n_interval - a.rec_origin : -1;
      p.it = it;
      p.rec_slot = (slot >= 0 && slot < a.rec_cap) ? slot : -1;
      p.trace = a.trace ? a.trace + static_cast<size_t>(it - a.iter_first) * per_it_s : nullptr;
      gibbs_sweep_kernel<real, TRACE><<<grid, kCW * 32, smem_iter, stream>>>(p, d, nullptr);
      gibbs_totals_kernel<<<(a.J + 3) / 4, 128, 0, stream>>>(p.part_cur, p.part_prop, a.J, p.n_tiles, a.totals);
      if (!nccl_allreduce_f64(a.comm, a.totals, a.totals, 2 * static_cast<size_t>(a.J), stream, &why))
        return cudaErrorUnknown;
      GibbsDecide dd = d;
      dd.n_tiles = 1;  // the "partials" are now the global per-chain sums
      dd.part_cur = a.totals;
      dd.part_prop = a.totals + a.J;
      dd.it = it;
      dd.it_next = it + 1;
      dd.rec_slot = p.rec_slot;
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
  if (want != 0 && !peer_xch && smem_pers <= 200 * 1024) {
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
  err = cudaFuncSetAttribute(gibbs_sweep_kernel<real, TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem_iter));
  if (err != cudaSuccess) return err;
  const size_t per_it = static_cast<size_t>(a.E + 1) * a.J;
  for (int it = a.iter_first; it <= a.iter_last; ++it) {
    const bool rec = a.n_interval > 1 && (it % a.n_interval) == 1;
    const int slot = rec ? (it - 1) / a.n_interval - a.rec_origin : -1;
    p.it = it;
    p.rec_slot = (slot >= 0 && slot < a.rec_cap) ? slot : -1;
    p.trace = a.trace ? a.trace + static_cast<size_t>(it - a.iter_first) * per_it : nullptr;
    d.it = it;
    d.it_next = it + 1;
    d.rec_slot = p.rec_slot;
    d.trace = a.trace ? a.trace + static_cast<size_t>(it - a.iter_first) * per_it + static_cast<size_t>(a.E) * a.J : nullptr;
    d.swap = a.swaps ? a.swaps + (it - a.iter_first) : nullptr;
    d.xch.epoch = a.xch_epoch0 + static_cast<uint32_t>(it - a.iter_first);
    gibbs_sweep_kernel<real, TRACE><<<grid, kCW * 32, smem_iter, stream>>>(p, d, a.done_counter);
    ++nl;
  }
  if (n_launches) *n_launches = nl;
  return cudaGetLastError();
}

cudaError_t launch_gibbs(const GibbsLaunch& a, cudaStream_t stream, int* n_launches) {
  if (a.precision == HTM_PRECISION_F32) return launch_gibbs_f32(a, stream, n_launches);
  return (a.trace || a.swaps) ? launch_gibbs_tt<double, true>(a, stream, n_launches)
                              : launch_gibbs_tt<double, false>(a, stream, n_launches);
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
  if (a.precision == HTM_PRECISION_F32) return launch_gibbs_f32_init(a, stream);
  const size_t n = static_cast<size_t>(a.J) * a.E;
  const unsigned grid = static_cast<unsigned>((n + 127) / 128);
  gibbs_init_hypo_kernel<double><<<grid, 128, 0, stream>>>(make_gibbs_params<double>(a), a.seed);
  gibbs_total_kernel<double><<<(a.J + 3) / 4, 128, 0, stream>>>(static_cast<const double*>(a.hLe), a.E, a.J, a.g_L);
  return cudaGetLastError();
}

}  // namespace htm
