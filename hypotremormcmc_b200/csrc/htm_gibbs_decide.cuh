// Mode C (blocked Gibbs): parameter blocks, chain-level state in shared memory and the decide step
// (sum of the partials, optional exchange between event shards, judge, record, swap, next proposal).
// Included by htm_gibbs.cu only.
#pragma once
#include "htm_forward.cuh"
#include "htm_kernels.hpp"

namespace htm {

constexpr int kTile = 32;   // events per CTA
constexpr int kCW = 8;      // chains (warps) per CTA

template <typename real>
struct GibbsParams {
  typedef typename M<real>::real4 real4;
  const real4* sta4;
  const real4* obs4;  // raw: no station terms folded
  const real4* evc4;
  const void* prior_xy;  // real2 [E]
  const float4* obsx;    // float32 only: expanded station-pair rows [E][xrow] (htm_gibbs_f32.cu)
  int xrow;
  real *hx, *hy, *hz, *hLe, *hLp;  // [J][E]
  double *g_vs, *g_qs, *g_tc, *g_ac, *g_T, *g_L;  // [J], [J][S]
  int* prop_which;
  int* prop_idx;
  double* prop_xnew;
  double* prop_lpr;
  int* a_prev;
  int* slot_of;
  double *part_cur, *part_prop;  // [J][n_tiles]
  int E, S, J, K, n_tiles, n_cool_total;
  int it, n_burn, n_interval;
  PhiloxKeys rk;
  uint32_t event_offset;
  uint32_t chain_offset, J_total;  // Philox ids are global: shards of virtual ranks draw distinct streams
  real prior_z, width_z, width_xy, step_xy, step_z;
  unsigned long long* counts;
  real4* hypo_rec;  // [cap][n_cool_total][E]
  int rec_slot;     // ring slot of this iteration, or -1
  htm_step_trace* trace;  // this iteration's [E+1][J] block, or null
  htm_swap_trace* swap;   // this iteration's record, or null
};

template <typename real>
__device__ __forceinline__ bool gibbs_is_cold(double T) {
  return T < 1.0 + kEps64;
}


// ---- chain-level bookkeeping -----------------------------------------------------------------------
struct GibbsDecide {
  double *g_vs, *g_qs, *g_tc, *g_ac, *g_T, *g_L;
  int* prop_which;
  int* prop_idx;
  double* prop_xnew;
  double* prop_lpr;
  int* a_prev;
  int* slot_of;
  const double *part_cur, *part_prop;
  int S, J, K, n_tiles, n_cool_total;
  int it;       // iteration being decided; 0 = prepare only (no decision, no swap)
  int it_next;  // iteration to propose for
  PhiloxKeys rk;
  uint32_t chain_offset, swap_stream;
  int n_solved;
  int solved[4];
  double prior[4], width[4], step[4];  // indexed by type-1: vs, t_corr, qs, a_corr
  unsigned long long* counts;
  int count_globals;  // 0 on shards > 0 of an event-sharded run (the decisions are replicated)
  PeerExchange xch;   // event shards: sums of all shards through peer memory (n <= 1: off)
  // shared-parameter records of the cold chains: [cap][n_cool_total]
  int rec_slot;
  int* rec_chain;
  double *rec_vs, *rec_qs, *rec_L, *rec_tc, *rec_ac;
  htm_step_trace* trace;  // [J] (row E of this iteration's block) or null
  htm_swap_trace* swap;
};

__device__ __forceinline__ double gauss64(uint32_t wa, uint32_t wb) { return M<double>::gauss(wa, wb); }

// Chain-level state staged in shared memory: the serial parts of the decide step (swap, cold-slot numbering)
// never wait on global memory, and the persistent kernel keeps it there for the whole launch.
struct ChainSm {
  double *T, *L, *vs, *qs, *xnew, *lpr, *tot, *tc, *ac;  // tot: [2][J]; tc, ac: [J][S]
  int *which, *idx, *aprev, *slot;
};
__host__ __device__ inline size_t chain_sm_bytes(int J, int S) {
  return static_cast<size_t>(J) * (2 * S + 8) * sizeof(double) + static_cast<size_t>(J) * 4 * sizeof(int);
}
__device__ __forceinline__ ChainSm carve_chain_sm(unsigned char* base, int J, int S) {
  ChainSm c;
  double* d = reinterpret_cast<double*>(base);
  c.T = d; d += J;
  c.L = d; d += J;
  c.vs = d; d += J;
  c.qs = d; d += J;
  c.xnew = d; d += J;
  c.lpr = d; d += J;
  c.tot = d; d += 2 * J;
  c.tc = d; d += static_cast<size_t>(J) * S;
  c.ac = d; d += static_cast<size_t>(J) * S;
  int* i = reinterpret_cast<int*>(d);
  c.which = i; i += J;
  c.idx = i; i += J;
  c.aprev = i; i += J;
  c.slot = i;
  return c;
}
// Variant for any number of joint chains: only the small per-chain state lives in shared memory; the station terms
// t_corr / a_corr [J][S] stay in global memory (L2) and cs.tc / cs.ac point there.  Every CTA that runs the
// decide step redundantly writes the SAME accepted values, and every read of a term bypasses L1 (term_ld), so
// no CTA can see a stale value (htm_gibbs_f32.cu).
__host__ __device__ inline size_t chain_sm_small_bytes(int J) {
  return static_cast<size_t>(J) * 8 * sizeof(double) + static_cast<size_t>(J) * 4 * sizeof(int);
}
__device__ __forceinline__ ChainSm carve_chain_sm_small(unsigned char* base, int J, double* g_tc, double* g_ac) {
  ChainSm c;
  double* d = reinterpret_cast<double*>(base);
  c.T = d; d += J;
  c.L = d; d += J;
  c.vs = d; d += J;
  c.qs = d; d += J;
  c.xnew = d; d += J;
  c.lpr = d; d += J;
  c.tot = d; d += 2 * J;
  c.tc = g_tc;
  c.ac = g_ac;
  int* i = reinterpret_cast<int*>(d);
  c.which = i; i += J;
  c.idx = i; i += J;
  c.aprev = i; i += J;
  c.slot = i;
  return c;
}
__device__ __forceinline__ double term_ld(const double* p, const bool gterms) { return gterms ? __ldcg(p) : *p; }

static __device__ void chain_load(const GibbsDecide& d, const ChainSm& cs, const bool terms = true) {
  for (int c = threadIdx.x; c < d.J; c += blockDim.x) {
    cs.T[c] = d.g_T[c];
    cs.L[c] = d.g_L[c];
    cs.vs[c] = d.g_vs[c];
    cs.qs[c] = d.g_qs[c];
    cs.xnew[c] = d.prop_xnew[c];
    cs.lpr[c] = d.prop_lpr[c];
    cs.which[c] = d.prop_which[c];
    cs.idx[c] = d.prop_idx[c];
    cs.aprev[c] = d.a_prev[c];
    cs.slot[c] = d.slot_of[c];
  }
  if (!terms) return;
  for (int i = threadIdx.x; i < d.J * d.S; i += blockDim.x) {
    cs.tc[i] = d.g_tc[i];
    cs.ac[i] = d.g_ac[i];
  }
}
static __device__ void chain_store(const GibbsDecide& d, const ChainSm& cs, const bool terms = true) {
  for (int c = threadIdx.x; c < d.J; c += blockDim.x) {
    d.g_T[c] = cs.T[c];
    d.g_L[c] = cs.L[c];
    d.g_vs[c] = cs.vs[c];
    d.g_qs[c] = cs.qs[c];
    d.prop_xnew[c] = cs.xnew[c];
    d.prop_lpr[c] = cs.lpr[c];
    d.prop_which[c] = cs.which[c];
    d.prop_idx[c] = cs.idx[c];
    d.a_prev[c] = cs.aprev[c];
    d.slot_of[c] = cs.slot[c];
  }
  if (!terms) return;
  for (int i = threadIdx.x; i < d.J * d.S; i += blockDim.x) {
    d.g_tc[i] = cs.tc[i];
    d.g_ac[i] = cs.ac[i];
  }
}

// ---- event shards: all-reduce of the per-chain sums over NVLink peer memory --------------------------------
// Called by every thread of ONE CTA per shard (the last CTA of the sweep).  Each shard stores its W sums into
// slot [parity][rank] of every shard's buffer, publishes the exchange number with a system-scope release, waits
// for the numbers of all shards, and adds the n slots of its own buffer in shard order -- the same operands in
// the same order everywhere, so every shard takes bit-identical decisions.  Two parities: a shard can be at most
// one exchange ahead of the slowest one (it cannot pass exchange e+1 before everyone has published e+1, i.e.
// has finished reading e).
// The wait is bounded by WALL CLOCK (globaltimer; shards are separate host processes that may drain gigabytes
// of samples between chunks, so a poll count would be a guess): x.timeout_ns per wait, minutes by default.
// Returns false -- uniformly for the CTA -- when a peer did not answer now or in an earlier exchange; *status
// is then set, the caller must take NO decision from `tot`, and every result-returning entry point of the C
// ABI reports the failure (htm_capi.cu: check_exchange).
__device__ __forceinline__ uint32_t* peer_flags(double* base, int n, int W) {
  return reinterpret_cast<uint32_t*>(base + static_cast<size_t>(2) * n * W);
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
static __device__ bool peer_allreduce(const PeerExchange& x, const uint32_t epoch, double* tot /* shared memory, W values, in/out */,
                               const int W) {
  __shared__ int s_xch_fail;
  const int n = x.n, me = x.rank, par = static_cast<int>(epoch & 1u);
  if (threadIdx.x == 0) s_xch_fail = (x.status && *reinterpret_cast<volatile int*>(x.status) != 0) ? 1 : 0;
  __syncthreads();
  if (s_xch_fail) return false;  // the run is lost already: neither publish nor wait
  for (int i = threadIdx.x; i < n * W; i += blockDim.x) {
    const int r = i / W, t = i - r * W;
    x.peer[r][(static_cast<size_t>(par) * n + me) * W + t] = tot[t];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < static_cast<unsigned>(n)) {
    uint32_t* theirs = peer_flags(x.peer[threadIdx.x], n, W) + par * n + me;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
    const uint32_t* mine = peer_flags(x.peer[me], n, W) + par * n + threadIdx.x;
    uint32_t seen = 0;
    const unsigned long long t0 = global_timer_ns();
    unsigned int polls = 0;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
      if (seen == epoch) break;
      if ((++polls & 1023u) == 0u && global_timer_ns() - t0 > x.timeout_ns) {
        if (x.status) *reinterpret_cast<volatile int*>(x.status) = 1;
        s_xch_fail = 1;
        break;
      }
    }
  }
  __syncthreads();
  if (s_xch_fail) {
    __threadfence();  // the flag is visible to the rest of the grid before anyone acts on it
    return false;
  }
  for (int t = threadIdx.x; t < W; t += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < n; ++r) s += __ldcg(x.peer[me] + (static_cast<size_t>(par) * n + r) * W + t);
    tot[t] = s;
  }
  __syncthreads();
  return true;
}

// Per-chain sums of the per-tile (or per-CTA) partial sums, in a fixed order; every thread of the CTA calls it
// (no barrier inside: the caller synchronises before reading tot).
static __device__ void sum_partials(const int n_tiles, const double* part_cur, const double* part_prop, const int J, double* tot) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  // fixed-order sums of the per-tile partials: lanes stride over tiles, then a butterfly.  With fewer than
  // 32 partials per chain a warp serves 32/g chains at once (g = lanes per chain, a power of two): the
  // butterfly levels it skips would only have added the zeros of lanes that hold no partial, so the sums
  // are the same bits as with one chain per warp.
  if (n_tiles <= 16) {
    int g = 16;
    while (g > 1 && (g >> 1) >= n_tiles) g >>= 1;
    const int per_warp = 32 / g, sub = lane / g, l = lane - sub * g;
    for (int c0 = warp * per_warp; c0 < J; c0 += nw * per_warp) {
      const int c = c0 + sub;
      double a = 0.0, b = 0.0;
      if (c < J && l < n_tiles) {  // g >= n_tiles: at most one partial per lane
        a = __ldcg(part_cur + static_cast<size_t>(c) * n_tiles + l);
        b = __ldcg(part_prop + static_cast<size_t>(c) * n_tiles + l);
      }
      for (int o = g >> 1; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if (l == 0 && c < J) {
        tot[c] = a;
        tot[J + c] = b;
      }
    }
  } else {
    // eight chains per pass, so that sixteen independent loads are in flight per lane instead of two (the step is a
    // chain of L2 round trips: with 100 chains x 74 partials it was most of the 9 us decide step)
    constexpr int kCh = 8;
    for (int c0 = warp * kCh; c0 < J; c0 += nw * kCh) {
      double a[kCh], b[kCh];
#pragma unroll
      for (int q = 0; q < kCh; ++q) a[q] = b[q] = 0.0;
      for (int t = lane; t < n_tiles; t += 32) {
#pragma unroll
        for (int q = 0; q < kCh; ++q) {
          const int c = min(c0 + q, J - 1);
          a[q] += __ldcg(part_cur + static_cast<size_t>(c) * n_tiles + t);
          b[q] += __ldcg(part_prop + static_cast<size_t>(c) * n_tiles + t);
        }
      }
#pragma unroll
      for (int q = 0; q < kCh; ++q) {
        const double sa = warp_sum<double>(a[q]), sb = warp_sum<double>(b[q]);
        if (lane == 0 && c0 + q < J) {
          tot[c0 + q] = sa;
          tot[J + c0 + q] = sb;
        }
      }
    }
  }
}

// The decide step on the staged state.  Every thread of the CTA takes part; a CTA that is not the
// `writer` computes exactly the same values but leaves counters, records and traces alone (the
// persistent kernel runs this redundantly on every CTA so that one grid barrier per iteration suffices).
// Ends with a block barrier.
static __device__ void decide_core(const GibbsDecide& d, const ChainSm& cs, const int it, const int it_next,
                            const double* part_cur, const double* part_prop, const int rec_slot,
                            htm_step_trace* trace, htm_swap_trace* swap, const bool writer,
                            const bool summed = false /* part_* are already the sums over all tiles and shards */,
                            const bool gterms = false /* cs.tc / cs.ac are global memory (carve_chain_sm_small) */) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int J = d.J, S = d.S;
  if (it > 0) {
    if (part_cur) {  // null: the caller has filled cs.tot (htm_gibbs_f32.cu: exact integer sums)
      sum_partials(summed ? 1 : d.n_tiles, part_cur, part_prop, J, cs.tot);
      __syncthreads();
    }
    // a failed exchange ends the run: no decision is taken from partial sums (uniform for the CTA)
    if (!summed && d.xch.n > 1 && !peer_allreduce(d.xch, d.xch.epoch, cs.tot, 2 * J)) return;
    // ---- judge the shared-parameter proposal (src/cls_mcmc.f90:186-219) ----
    for (int c = threadIdx.x; c < J; c += blockDim.x) {
      const int which = cs.which[c];
      const double T = cs.T[c];
      const bool cold = T < 1.0 + kEps64;
      const double Lcur = cs.tot[c], Lprop = cs.tot[J + c];
      bool acc = false;
      if (which != 0) {
        const u32x4 wb = philox4x32_10(d.rk, static_cast<uint32_t>(it), d.chain_offset + static_cast<uint32_t>(c), PHX_GLOBAL, 1u);
        const double ratio = (Lprop - Lcur) / T + cs.lpr[c];
        const double r = M<double>::u_co(wb.v[0]);
        if (r >= kEps64 && ::log(r) <= ratio) acc = true;
        if (writer && cold && d.counts && d.count_globals) {
          atomicAdd(d.counts + (which - 1), 1ull);
          if (acc) atomicAdd(d.counts + 7 + (which - 1), 1ull);
        }
        if (acc) {
          const double xn = cs.xnew[c];
          const int idx = cs.idx[c];
          if (which == 1) cs.vs[c] = xn;
          if (which == 2) cs.tc[static_cast<size_t>(c) * S + idx] = xn;
          if (which == 3) cs.qs[c] = xn;
          if (which == 4) cs.ac[static_cast<size_t>(c) * S + idx] = xn;
        }
      }
      cs.aprev[c] = acc ? 1 : 0;
      cs.L[c] = acc ? Lprop : Lcur;
      if (writer && trace) {
        htm_step_trace t;
        t.proposal_type = which;
        t.index = which ? cs.idx[c] + 1 : 0;
        t.prior_ok = 1;
        t.accepted = acc ? 1 : 0;
        t.log_likelihood = cs.L[c];
        trace[c] = t;
      }
    }
    __syncthreads();
    // ---- record the cold chains' shared parameters (src/hypo_tremor_mcmc.f90:270-280) ----
    if (writer && rec_slot >= 0 && d.rec_chain) {
      for (int c = warp; c < J; c += nw) {
        const int sl = cs.slot[c];
        if (sl < 0) continue;
        const size_t o = static_cast<size_t>(rec_slot) * d.n_cool_total + sl;
        if (lane == 0) {
          d.rec_chain[o] = c;
          d.rec_vs[o] = cs.vs[c];
          d.rec_qs[o] = cs.qs[c];
          d.rec_L[o] = cs.L[c];
        }
        for (int j = lane; j < S; j += 32) {
          d.rec_tc[o * S + j] = term_ld(cs.tc + static_cast<size_t>(c) * S + j, gterms);
          d.rec_ac[o * S + j] = term_ld(cs.ac + static_cast<size_t>(c) * S + j, gterms);
        }
      }
    }
    // ---- one swap attempt over all J chains (src/cls_parallel.f90:220-240, 285-302) ----
    if (threadIdx.x == 0 && J >= 2) {
      const u32x4 w = philox4x32_10(d.rk, static_cast<uint32_t>(it), d.swap_stream, PHX_SWAP, 1u);
      const int i1 = static_cast<int>(below(w.v[0], static_cast<uint32_t>(J)));
      int i2 = i1 + 1 + static_cast<int>(below(w.v[1], static_cast<uint32_t>(J - 1)));
      if (i2 >= J) i2 -= J;
      const double T1 = cs.T[i1], T2 = cs.T[i2], L1 = cs.L[i1], L2 = cs.L[i2];
      const double del_s = (L2 - L1) * (1.0 / T1 - 1.0 / T2);
      const double r = M<double>::u_co(w.v[2]);
      const bool sacc = r >= kEps64 && ::log(r) <= del_s;
      if (sacc) {
        cs.T[i1] = T2;
        cs.T[i2] = T1;
      }
      if (writer && swap) {
        htm_swap_trace t;
        t.rank1 = i1 / d.K;
        t.chain1 = i1 % d.K + 1;
        t.rank2 = i2 / d.K;
        t.chain2 = i2 % d.K + 1;
        t.accepted = sacc ? 1 : 0;
        t.reserved = 0;
        *swap = t;
      }
    }
    __syncthreads();
  }
  // ---- slots of the cold chains (in chain order) for the next iteration's records ----
  {
    __shared__ int s_cold_in_warp[32];
    int before = 0;  // cold chains in earlier passes
    for (int base = 0; base < J; base += blockDim.x) {
      const int c = base + threadIdx.x;
      const bool cold = c < J && cs.T[c] < 1.0 + kEps64;
      const uint32_t m = __ballot_sync(0xffffffffu, cold);
      if (lane == 0) s_cold_in_warp[warp] = __popc(m);
      __syncthreads();
      int off = before, all = 0;
      for (int w = 0; w < nw; ++w) {
        const int n = s_cold_in_warp[w];
        off += w < warp ? n : 0;
        all += n;
      }
      if (c < J) cs.slot[c] = cold ? off + __popc(m & ((1u << lane) - 1u)) : -1;
      before += all;
      __syncthreads();
    }
  }
  // ---- next shared-parameter proposal (src/cls_mcmc.f90:134-157 restricted to the solved ones) ----
  for (int c = threadIdx.x; c < J; c += blockDim.x) {
    if (d.n_solved == 0) {
      cs.which[c] = 0;
      continue;
    }
    const u32x4 wa = philox4x32_10(d.rk, static_cast<uint32_t>(it_next), d.chain_offset + static_cast<uint32_t>(c), PHX_GLOBAL, 0u);
    const int which = d.solved[below(wa.v[0], static_cast<uint32_t>(d.n_solved))];
    const int idx = (which == 2 || which == 4) ? static_cast<int>(below(wa.v[1], static_cast<uint32_t>(S))) : 0;
    const double gs = gauss64(wa.v[2], wa.v[3]);
    double x_old;
    if (which == 1) x_old = cs.vs[c];
    else if (which == 2) x_old = term_ld(cs.tc + static_cast<size_t>(c) * S + idx, gterms);
    else if (which == 3) x_old = cs.qs[c];
    else x_old = term_ld(cs.ac + static_cast<size_t>(c) * S + idx, gterms);
    const double mu = d.prior[which - 1], sg = d.width[which - 1];
    const double x_new = __dadd_rn(x_old, __dmul_rn(gs, d.step[which - 1]));
    const double dn = x_new - mu, dl = x_old - mu;
    cs.which[c] = which;
    cs.idx[c] = idx;
    cs.xnew[c] = x_new;
    cs.lpr[c] = -(__dmul_rn(dn, dn) - __dmul_rn(dl, dl)) / (2.0 * sg * sg);
  }
  __syncthreads();
}

// one CTA, station terms left in global memory: smem chain_sm_small_bytes(J)
static __device__ void gibbs_decide_small(const GibbsDecide& d, unsigned char* smem) {
  const ChainSm cs = carve_chain_sm_small(smem, d.J, d.g_tc, d.g_ac);
  chain_load(d, cs, false);
  __syncthreads();
  decide_core(d, cs, d.it, d.it_next, d.part_cur, d.part_prop, d.rec_slot, d.trace, d.swap, true, false, true);
  chain_store(d, cs, false);
}

// one CTA: global -> shared, decide, shared -> global.  smem: chain_sm_bytes(J, S)
static __device__ void gibbs_decide(const GibbsDecide& d, unsigned char* smem) {
  const ChainSm cs = carve_chain_sm(smem, d.J, d.S);
  chain_load(d, cs);
  __syncthreads();
  decide_core(d, cs, d.it, d.it_next, d.part_cur, d.part_prop, d.rec_slot, d.trace, d.swap, true);
  chain_store(d, cs);
}

// ---- host side: launch description -> kernel parameter blocks (shared by htm_gibbs.cu and htm_gibbs_f32.cu) ----
template <typename real>
inline GibbsParams<real> make_gibbs_params(const GibbsLaunch& a) {
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

inline GibbsDecide make_decide(const GibbsLaunch& a) {
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


// float32 joint-chain kernel (htm_gibbs_f32.cu)
cudaError_t launch_gibbs_f32(const GibbsLaunch& a, cudaStream_t stream, int* n_launches);
cudaError_t launch_gibbs_f32_init(const GibbsLaunch& a, cudaStream_t stream);

}  // namespace htm
