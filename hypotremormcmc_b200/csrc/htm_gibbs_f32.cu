// Mode C (blocked Gibbs), float32: the throughput kernel of the joint-chain schedule.
//
// Schedule (identical to the float64 kernels of htm_gibbs.cu, whose CPU statement lives with the test oracle):
// per iteration every (chain, event) proposes one hypocentre coordinate, judged on the event's own
// log-likelihood with the chain's temperature (src/cls_mcmc.f90:159-165,193-203; src/cls_forward.f90:307-362);
// then ONE shared parameter per chain (vs | one t_corr | qs | one a_corr) is judged on the sum over ALL events
// (what the reference recomputes with forward%calc_log_likelihood, src/cls_forward.f90:268-303); record; one
// swap attempt over all chains (src/cls_parallel.f90:220-240,285-302).
//
// What is different from a literal evaluation: the effect of a shared-parameter proposal on an event's
// log-likelihood is a LOW-ORDER POLYNOMIAL in the parameter change once a few weighted moments of the event's
// residuals are known, so the second O(n_sta) forward evaluation per (chain, event, iteration) disappears:
//   with e_j = r_j - r_0 (residual relative to station 0, j >= 1), D_j = d_j - d_0, w = sigma^-2,
//     S1 = sum w e      S2 = sum w e^2      A1 = sum w D      A2 = sum w e D      A3 = sum w D^2
//     chi^2 = S2 - S1^2 / W,  L_e = -(chi_t^2 + chi_a^2)/2 - C_e            (src/cls_forward.f90:166-173,283-297)
//   vs -> vs'    : e_t += D (1/vs' - 1/vs),  e_a -= D (B' - B),  B = pi f /(Q vs)      (:157-160, :247)
//                  dS1 = q A1, dS2 = q (2 A2 + q A3), dA2 = q A3      (q = the coefficient of D)
//   qs -> qs'    : the amplitude half of the same
//   tc_j -> +dlt : j >= 1: e_j -= dlt      -> dS1 = -w dlt, dS2 = w dlt (dlt - 2 e_j), dA2 = -w dlt D_j
//                  j == 0: e_k += dlt (all k) -> dS1 = dlt W', dS2 = dlt (2 S1 + dlt W'), dA2 = dlt A1
//   ac_j         : the same on the amplitude sums
// One pass over the stations at the PROPOSED hypocentre yields L_e and all moments (25 packed FP32 operations +
// 4 MUFU per station pair and chain instead of 2 x (16 + 4)); the per-event change of the pending
// shared-parameter proposal then costs O(1).  It is also the cancellation-free way to form the Metropolis ratio
// of a shared parameter in float32: sum_e dL_e is accumulated in float64 from per-event DIFFERENCES instead of
// differencing two sums of 10^4..10^5 float32 log-likelihoods.
//
// State per (chain, event), float4 arrays [J][E] (event index fastest):
//   H = {x, y, z, L_e}   M = {A1t, A3t, A1a, A3a}   Q = {S1t, S1a, A2t, A2a}   P = Q under the pending proposal
//   Lp = L_e under the pending proposal.  An accepted shared-parameter proposal is committed lazily by the next
//   sweep (H.w := Lp, Q := P), exactly the numbers that were judged.
//
// Mapping: persistent cooperative kernel, one wave of CTAs.  warp = 8 events x 4 chains (chain minor: lanes that
// share an event row are neighbours, 2 shared-memory wavefronts per 16-byte row read instead of 4); a CTA owns up
// to 32 chains and walks a contiguous range of event octets whose expanded rows (htm_forward.cuh) arrive through
// a 2-stage ring of 1-D bulk-TMA copies (full / empty mbarriers) that runs across iteration boundaries.  Per
// iteration: sweep -> per-(chain, CTA) float64 partial sums -> ONE grid barrier -> every CTA adds the partials in
// the same fixed order and takes the same decisions on its own shared-memory copy of the small per-chain state
// (CTA (0,0) alone writes counters, records, traces).  The station terms t_corr/a_corr [J][S] stay in global
// memory (L2): every CTA writes the same accepted values, reads bypass L1.  Event shards (several GPUs, one
// ensemble) exchange the per-chain sums through NVLink peer memory between two grid barriers.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>

#include "htm_gibbs_decide.cuh"

namespace cg = cooperative_groups;

namespace htm {

// tuning macros (tools/build_variants.sh): the register budget per thread (it decides how many CTAs fit an SM), and whether the
// per-lane state of the next octet visit is prefetched into shared memory with cp.async
#ifndef HTM_GIBBS_MAXREG
#define HTM_GIBBS_MAXREG 128
#endif
#ifndef HTM_GIBBS_PREFETCH
#define HTM_GIBBS_PREFETCH 1
#endif
constexpr bool kPrefetch = HTM_GIBBS_PREFETCH != 0;
#ifndef HTM_GIBBS_STAGES
#define HTM_GIBBS_STAGES 2
#endif
constexpr int kMaxStages = HTM_GIBBS_STAGES;  // depth of the TMA ring of event-octet rows (2 when shared memory is short)
constexpr int kOct = 8;
constexpr int kQuad = 4;
constexpr float kHalfLn2 = 0.34657359027997264f;

// ---- expanded rows: built once per table upload -------------------------------------------------------------
// row[0] = A of station 0, row[1] = {t_obs0, a_obs0, sum_{j>=1} w_t, sum_{j>=1} w_a}, row[2] = the event's
// constants {C_e, 1/W_t, 1/W_a, 0}, row[3] = {x_mu, y_mu, 0, 0} (so that everything an event needs arrives with
// its ONE bulk copy), then 4 float4 per station pair (1,2), (3,4), ... (store_station_pair); an even station
// count leaves a zero-weight half pair.
constexpr int kRowHead = 4;
__host__ __device__ inline int f32_xrow(int S) { return kRowHead + 4 * (S / 2); }
__global__ void expand_obs_kernel(const float4* __restrict__ sta4, const float4* __restrict__ obs4,
                                  const float4* __restrict__ evc4, const float2* __restrict__ prior_xy, int E, int S,
                                  float4* __restrict__ obsx) {
  const int n_pairs = S / 2, per_ev = n_pairs + 1, xrow = f32_xrow(S);
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(E) * per_ev) return;
  const int e = static_cast<int>(i / per_ev), m = static_cast<int>(i % per_ev) - 1;
  const float2 c = prior_xy[e];
  const float4* ob = obs4 + static_cast<size_t>(e) * S;
  float4* row = obsx + static_cast<size_t>(e) * xrow;
  if (m < 0) {
    const StaRecF r0 = expand_station(sta4[0], ob[0], c.x, c.y);
    float wt = 0.f, wa = 0.f;
    for (int j = 1; j < S; ++j) {
      wt += ob[j].y;
      wa += ob[j].w;
    }
    row[0] = r0.A;
    row[1] = make_float4(ob[0].x, ob[0].z, wt, wa);
    row[2] = evc4[e];
    row[3] = make_float4(c.x, c.y, 0.f, 0.f);
    return;
  }
  const int j0 = 1 + 2 * m, j1 = j0 + 1;
  const StaRecF a = expand_station(sta4[j0], ob[j0], c.x, c.y);
  StaRecF b = a;
  if (j1 < S)
    b = expand_station(sta4[j1], ob[j1], c.x, c.y);
  else
    b.B = make_float4(0.f, 0.f, 0.f, 0.f);
  store_station_pair(row + kRowHead + 4 * m, a, b);
}

cudaError_t launch_expand_obs(const Tables& tab, int E, int S, void* obsx, cudaStream_t stream) {
  const size_t n = static_cast<size_t>(E) * (S / 2 + 1);
  expand_obs_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, stream>>>(
      static_cast<const float4*>(tab.sta4), static_cast<const float4*>(tab.obs4_raw),
      static_cast<const float4*>(tab.evc4), static_cast<const float2*>(tab.prior_xy), E, S, static_cast<float4*>(obsx));
  return cudaGetLastError();
}

// ---- one pass over the stations: log-likelihood and moments at a hypocentre ---------------------------------
struct Moments {
  float L, S1t, S1a, A1t, A2t, A3t, A1a, A2a, A3a;
};
struct RowHead {  // station 0 at a hypocentre: its distance and the (negated) shifts
  float d0, nct, nca;
};
__device__ __forceinline__ RowHead row_head(const float4* __restrict__ row, const float hx, const float hy, const float hz,
                                            const float h2, const Glob<float>& g, const float ntc0, const float nac0) {
  const float4 A0 = row[0], h1 = row[1];
  const float d2 = fmaf(hx, A0.x, fmaf(hy, A0.y, fmaf(hz, A0.z, A0.w + h2)));
  RowHead r;
  r.d0 = d2 * mufu_rsq(d2);
  const float l2 = mufu_lg2(d2);
  r.nct = -(fmaf(r.d0, g.ivs, ntc0) - h1.x);
  r.nca = -(fmaf(-kHalfLn2, l2, fmaf(-g.B, r.d0, nac0)) - h1.y);
  return r;
}

#ifndef HTM_GIBBS_PK_UNROLL
#define HTM_GIBBS_PK_UNROLL 4  // 2: 105.5 / 929 us per iteration at 10 000 / 100 000 events x 50 x 100 chains, 4: 100.3 / 875, 6-12: 102 / 892-905 (profiles/r2bm_gibbs_unroll.txt)
#endif
constexpr int kMomentsUnroll = HTM_GIBBS_PK_UNROLL;
// cp[m] = {-tc_j0, -tc_j1, -ac_j0, -ac_j1}: the chain's station terms of pair m; ntc0 / nac0 those of station 0
__device__ __forceinline__ Moments eval_moments_f32(const float4* __restrict__ row, const int n_pairs, const float hx,
                                                    const float hy, const float hz, const Glob<float>& g,
                                                    const float4* __restrict__ cp, const float ntc0, const float nac0,
                                                    const float4 evc) {
  const float h2 = fmaf(hz, hz, fmaf(hy, hy, hx * hx));
  const RowHead hd = row_head(row, hx, hy, hz, h2, g, ntc0, nac0);
  const float2 nct = f2(hd.nct, hd.nct), nca = f2(hd.nca, hd.nca), nd0 = f2(-hd.d0, -hd.d0);
  const float2 px = f2(hx, hx), py = f2(hy, hy), pz = f2(hz, hz), hh = f2(h2, h2);
  const float2 ivs2 = f2(g.ivs, g.ivs), nB2 = f2(-g.B, -g.B), nc2 = f2(-kHalfLn2, -kHalfLn2);
  float2 a1t = f2(0.f, 0.f), a1a = f2(0.f, 0.f), a2 = f2(0.f, 0.f);
  float2 m1t = f2(0.f, 0.f), m2t = f2(0.f, 0.f), m3t = f2(0.f, 0.f);
  float2 m1a = f2(0.f, 0.f), m2a = f2(0.f, 0.f), m3a = f2(0.f, 0.f);
  const float4* r = row + kRowHead;
#pragma unroll kMomentsUnroll
  for (int m = 0; m < n_pairs; ++m) {
    const float4 r0 = r[4 * m], r1 = r[4 * m + 1], r2 = r[4 * m + 2], r3 = r[4 * m + 3];
    const float4 c4 = cp[m];
    const float2 swt = f2(r2.x, r2.y), swa = f2(r3.x, r3.y);
    const float2 d2 = __ffma2_rn(px, f2(r0.x, r0.y),
                                 __ffma2_rn(py, f2(r0.z, r0.w), __ffma2_rn(pz, f2(r1.x, r1.y), __fadd2_rn(f2(r1.z, r1.w), hh))));
    const float2 d = __fmul2_rn(d2, f2(mufu_rsq(d2.x), mufu_rsq(d2.y)));
    const float2 l2 = f2(mufu_lg2(d2.x), mufu_lg2(d2.y));
    const float2 at = __fadd2_rn(__ffma2_rn(d, ivs2, nct), f2(c4.x, c4.y));
    const float2 ut = __ffma2_rn(swt, at, f2(r2.z, r2.w));  // sqrt(w_t) e_t
    const float2 aa = __fadd2_rn(__ffma2_rn(nc2, l2, __ffma2_rn(nB2, d, nca)), f2(c4.z, c4.w));
    const float2 ua = __ffma2_rn(swa, aa, f2(r3.z, r3.w));  // sqrt(w_a) e_a
    const float2 D = __fadd2_rn(d, nd0);
    const float2 pt = __fmul2_rn(swt, D), pa = __fmul2_rn(swa, D);  // sqrt(w) D
    a2 = __ffma2_rn(ut, ut, a2);
    a1t = __ffma2_rn(swt, ut, a1t);
    a2 = __ffma2_rn(ua, ua, a2);
    a1a = __ffma2_rn(swa, ua, a1a);
    m1t = __ffma2_rn(swt, pt, m1t);
    m2t = __ffma2_rn(ut, pt, m2t);
    m3t = __ffma2_rn(pt, pt, m3t);
    m1a = __ffma2_rn(swa, pa, m1a);
    m2a = __ffma2_rn(ua, pa, m2a);
    m3a = __ffma2_rn(pa, pa, m3a);
  }
  Moments o;
  o.S1t = a1t.x + a1t.y;
  o.S1a = a1a.x + a1a.y;
  o.A1t = m1t.x + m1t.y;
  o.A2t = m2t.x + m2t.y;
  o.A3t = m3t.x + m3t.y;
  o.A1a = m1a.x + m1a.y;
  o.A2a = m2a.x + m2a.y;
  o.A3a = m3a.x + m3a.y;
  o.L = finish_loglik<float>(o.S1t, a2.x + a2.y, o.S1a, 0.f, evc);
  return o;
}

// Change of the sums under the chain's pending shared-parameter proposal (see the header of this file).
// qa: coefficient of D in e_t (vs: 1/vs' - 1/vs, else 0); qb: coefficient of D in e_a (vs, qs: -(B' - B));
// dlt: change of the station term (t_corr / a_corr proposals).
struct PropF32 {
  int which, idx;
  float qa, qb, dlt;
};
struct DeltaSums {
  float dS1t, dS2t, dA2t, dS1a, dS2a, dA2a;
};
// The four chains of a warp usually hold proposals of different kinds, so the station-term case is written
// WITHOUT divergent branches: every lane evaluates station 0 and "its" station at the current hypocentre (one
// gather from its own event row) and the results are selected by kind.  `any_station` is warp-uniform.
__device__ __forceinline__ DeltaSums pending_delta(const PropF32& pr, const bool any_station, const float4* __restrict__ row,
                                                   const float hx, const float hy, const float hz, const Glob<float>& g,
                                                   const float4* __restrict__ cp, const float ntc0, const float nac0,
                                                   const float4 Mm, const float4 Q) {
  DeltaSums o;
  // vs / qs: polynomial in the coefficients (qa = qb = 0 for the other kinds)
  o.dS1t = pr.qa * Mm.x;
  o.dS2t = pr.qa * fmaf(pr.qa, Mm.y, 2.f * Q.z);
  o.dA2t = pr.qa * Mm.y;
  o.dS1a = pr.qb * Mm.z;
  o.dS2a = pr.qb * fmaf(pr.qb, Mm.w, 2.f * Q.w);
  o.dA2a = pr.qb * Mm.w;
  if (!any_station) return o;
  const bool is_t = pr.which == 2, is_a = pr.which == 4;
  const float dlt = pr.dlt;
  const float h2 = fmaf(hz, hz, fmaf(hy, hy, hx * hx));
  const RowHead hd = row_head(row, hx, hy, hz, h2, g, ntc0, nac0);
  const int j = max(pr.idx, 1) - 1, m = j >> 1;
  const bool hi = (j & 1) != 0;
  const float4* r = row + kRowHead + 4 * m;
  const float4 r0 = r[0], r1 = r[1], r2 = r[2], r3 = r[3], c4 = cp[m];
  const float cx = hi ? r0.y : r0.x, cy = hi ? r0.w : r0.z, cz = hi ? r1.y : r1.x, cc = hi ? r1.w : r1.z;
  const float d2 = fmaf(hx, cx, fmaf(hy, cy, fmaf(hz, cz, cc + h2)));
  const float d = d2 * mufu_rsq(d2);
  const float l2 = mufu_lg2(d2);
  const float D = d - hd.d0;
  // station j >= 1 of kind t / a: e_j -= dlt
  const float swt = hi ? r2.y : r2.x, sot = hi ? r2.w : r2.z, ntc = hi ? c4.y : c4.x;
  const float swa = hi ? r3.y : r3.x, soa = hi ? r3.w : r3.z, nac = hi ? c4.w : c4.z;
  const float ut = fmaf(swt, fmaf(d, g.ivs, hd.nct) + ntc, sot);  // sqrt(w) e
  const float ua = fmaf(swa, fmaf(-kHalfLn2, l2, fmaf(-g.B, d, hd.nca)) + nac, soa);
  const float wt = swt * swt, wa = swa * swa;
  float s1t = -wt * dlt, s2t = dlt * fmaf(wt, dlt, -2.f * swt * ut), a2t = -wt * dlt * D;
  float s1a = -wa * dlt, s2a = dlt * fmaf(wa, dlt, -2.f * swa * ua), a2a = -wa * dlt * D;
  // station 0 carries the shift: every other residual moves by +dlt
  const float4 h1 = row[1];
  const bool first = pr.idx == 0;
  s1t = first ? dlt * h1.z : s1t;
  s2t = first ? dlt * fmaf(dlt, h1.z, 2.f * Q.x) : s2t;
  a2t = first ? dlt * Mm.x : a2t;
  s1a = first ? dlt * h1.w : s1a;
  s2a = first ? dlt * fmaf(dlt, h1.w, 2.f * Q.y) : s2a;
  a2a = first ? dlt * Mm.z : a2a;
  o.dS1t = is_t ? s1t : o.dS1t;
  o.dS2t = is_t ? s2t : o.dS2t;
  o.dA2t = is_t ? a2t : o.dA2t;
  o.dS1a = is_a ? s1a : o.dS1a;
  o.dS2a = is_a ? s2a : o.dS2a;
  o.dA2a = is_a ? a2a : o.dA2a;
  return o;
}

// ---- shared memory of one CTA ----------------------------------------------------------------------------------
struct F32Sm {
  uint64_t* full;   // [n_stages]
  uint64_t* empty;  // [n_stages]
  float4* rows;     // [n_stages][kOct][row]
  float4* cp;       // [nc][cps]  {-tc_j0, -tc_j1, -ac_j0, -ac_j1}: station terms of the CTA's chains (persistent)
  float4* c0;       // [nc]  {-tc0, -ac0, 0, 0}
  float4* pf;       // [2][n_warps * 32][4]  per-lane state of the next octet visit (cp.async): H, M, Q | P, Lp
  float* pq;        // [nc][3]  qa, qb, dlt of the pending proposal
  int row, cps;
};
__host__ __device__ inline int f32_cps(int S) { return (S / 2) | 1; }
__host__ __device__ inline size_t f32_sweep_smem(int S, int nc, int n_stages) {
  return 16 * n_stages + (static_cast<size_t>(n_stages) * kOct * (f32_xrow(S) + 1) + static_cast<size_t>(nc) * (f32_cps(S) + 1) +
               (kPrefetch ? static_cast<size_t>(2) * (nc / kQuad) * 32 * 4 : 0)) * sizeof(float4) +
         static_cast<size_t>(nc) * 4 * sizeof(float);
}
__device__ __forceinline__ F32Sm carve_f32_sm(unsigned char* base, int S, int xrow, int nc, int n_stages) {
  F32Sm m;
  m.row = xrow + 1;
  m.cps = f32_cps(S);
  m.full = reinterpret_cast<uint64_t*>(base);
  m.empty = m.full + n_stages;
  m.rows = reinterpret_cast<float4*>(base + 16 * n_stages);
  m.cp = m.rows + n_stages * kOct * m.row;
  m.c0 = m.cp + nc * m.cps;
  m.pf = m.c0 + nc;
  m.pq = reinterpret_cast<float*>(m.pf + (kPrefetch ? 2 * (nc / kQuad) * 32 * 4 : 0));
  return m;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
               "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
               "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// The station terms of the CTA's chains from global memory (L2), once per launch; afterwards only the entry an
// accepted t_corr / a_corr proposal changed is rewritten (f32_update_chain_terms).
__device__ __forceinline__ void f32_stage_chain_terms(const F32Sm& m, const ChainSm& cs, int nc, int c_base, int J /* first chain past the CTA's */, int S) {
  const int n_pairs = S / 2;
  for (int i = threadIdx.x; i < nc * n_pairs; i += blockDim.x) {
    const int lc = i / n_pairs, mm = i - lc * n_pairs, c = c_base + lc;
    if (c >= J) continue;
    const double* gtc = cs.tc + static_cast<size_t>(c) * S;
    const double* gac = cs.ac + static_cast<size_t>(c) * S;
    const int j0 = 1 + 2 * mm, j1 = j0 + 1;
    float4 cur = make_float4(-static_cast<float>(__ldcg(gtc + j0)), 0.f, -static_cast<float>(__ldcg(gac + j0)), 0.f);
    if (j1 < S) {
      cur.y = -static_cast<float>(__ldcg(gtc + j1));
      cur.w = -static_cast<float>(__ldcg(gac + j1));
    }
    m.cp[lc * m.cps + mm] = cur;
  }
  for (int lc = threadIdx.x; lc < nc; lc += blockDim.x) {
    const int c = c_base + lc;
    if (c >= J) continue;
    m.c0[lc] = make_float4(-static_cast<float>(__ldcg(cs.tc + static_cast<size_t>(c) * S)),
                           -static_cast<float>(__ldcg(cs.ac + static_cast<size_t>(c) * S)), 0.f, 0.f);
  }
}
// one thread per chain of the CTA: the entry an accepted station-term proposal (which, idx, x_new) changed
__device__ __forceinline__ void f32_update_chain_term(const F32Sm& m, int lc, int which, int idx, double x_new) {
  const float v = -static_cast<float>(x_new);
  if (idx == 0) {
    if (which == 2) m.c0[lc].x = v;
    if (which == 4) m.c0[lc].y = v;
    return;
  }
  float* w = reinterpret_cast<float*>(m.cp + lc * m.cps + ((idx - 1) >> 1));
  w[(which == 2 ? 0 : 2) + ((idx - 1) & 1)] = v;
}
// one thread per chain of the CTA: coefficients of the chain's pending proposal
__device__ __forceinline__ void f32_stage_proposal(const F32Sm& m, const ChainSm& cs, int lc, int c, int S) {
  const int wh = cs.which[c];
  const double vs = cs.vs[c], qs = cs.qs[c], xn = cs.xnew[c];
  double qa = 0.0, qb = 0.0, dlt = 0.0;
  if (wh == 1) {
    qa = 1.0 / xn - 1.0 / vs;
    qb = -(kPi * kFreq / (qs * xn) - kPi * kFreq / (qs * vs));
  } else if (wh == 3) {
    qb = -(kPi * kFreq / (xn * vs) - kPi * kFreq / (qs * vs));
  } else if (wh == 2) {
    dlt = xn - __ldcg(cs.tc + static_cast<size_t>(c) * S + cs.idx[c]);
  } else if (wh == 4) {
    dlt = xn - __ldcg(cs.ac + static_cast<size_t>(c) * S + cs.idx[c]);
  }
  m.pq[3 * lc] = static_cast<float>(qa);
  m.pq[3 * lc + 1] = static_cast<float>(qb);
  m.pq[3 * lc + 2] = static_cast<float>(dlt);
}

struct F32State {
  float4 *H, *M, *Q, *P;
  float* Lp;
};

// -DHTM_GIBBS_PHASE_TRACE (tools/build_variants.sh + tools/gibbs_phase_trace.py): CTA (0,0) stamps globaltimer at the
// phase boundaries of the first 4096 iterations of a launch -- where an iteration's time goes (sweep, grid barrier,
// exchange, second barrier, decide).  Not compiled into the product library.
#ifdef HTM_GIBBS_PHASE_TRACE
__device__ unsigned long long g_phase_ns[8 * 4096];
__device__ unsigned long long g_cta_done_ns[2 * 4096];  // per CTA: sweep start / end of launch-relative iteration 10
__device__ unsigned int g_cta_info[4 * 4096];           // per CTA: SM id, blockIdx.y, event octets, active warps
#define HTM_PHASE(k)                                                                                  \
  do {                                                                                                \
    if (writer && threadIdx.x == 0 && !INIT && it - iter_first < 4096) g_phase_ns[(it - iter_first) * 8 + (k)] = global_timer_ns(); \
  } while (0)
#else
#define HTM_PHASE(k) \
  do {               \
  } while (0)
#endif

// INIT = true: generate_model for every (chain, event) (src/cls_model.f90:139-158, Philox draws as in the
// float64 path) and its state at the initial shared parameters; no iteration is run.
// ---- sums over CTAs (and GPUs) ------------------------------------------------------------------------------------
// A warp adds its lanes' log-likelihoods in float64 in a fixed order, as before; what a (warp, chain) contributes to
// the chain's sum over all events is then added to global memory as an INTEGER (fixed point, 2^-32 resolution, two
// 64-bit limbs, `red.add.u64`): integer adds commute, so the total does not depend on the order in which CTAs -- or
// event shards on other GPUs -- arrive, the result stays bitwise reproducible, and the decide step reads 4 J words
// instead of every CTA's partial sums (9.2 -> 4.1 us at 100 chains x 296 CTAs, 14.6 -> 4.6 us at 20 chains x 444).
// A contribution is rounded to 2^-32 (a float32 log-likelihood of magnitude >= 2^-8 has no bits below that).
constexpr double kFixScale = 4294967296.0;  // 2^32
constexpr int kLimbBits = 48;               // low limb: 48 bits, up to 2^16 contributions per word without overflow
__device__ __forceinline__ void publish_sum(unsigned long long* dst /* 2 limbs */, const double v) {
  const double hi = floor(v * 0x1p-16);                      // units of 2^48 * 2^-32
  const double lo = fma(-hi, 0x1p48, v * kFixScale);         // exact: the low bits of v 2^32, in [0, 2^48)
  atomicAdd(dst, __double2ull_rn(lo));
  atomicAdd(dst + 1, static_cast<unsigned long long>(__double2ll_rn(hi)));
}
__host__ __device__ inline double limbs_to_double(const unsigned long long l0, const unsigned long long l1) {
  const long long hi = static_cast<long long>(l1) + static_cast<long long>(l0 >> kLimbBits);
  const unsigned long long lo = l0 & ((1ull << kLimbBits) - 1);
  return (static_cast<double>(hi) * 281474976710656.0 /* 2^48 */ + static_cast<double>(lo)) * (1.0 / kFixScale);
}

// The sums of every shard's limbs, through peer memory (integer adds: the same bits on every shard whatever the
// shard order).  in / out: global memory, W words.  See peer_allreduce (htm_gibbs_decide.cuh) for the protocol.
static __device__ bool peer_allreduce_u64(const PeerExchange& x, const uint32_t epoch, const unsigned long long* in,
                                          unsigned long long* out, const int W) {
  __shared__ int s_xch_fail64;
  const int n = x.n, me = x.rank, par = static_cast<int>(epoch & 1u);
  if (threadIdx.x == 0) s_xch_fail64 = (x.status && *reinterpret_cast<volatile int*>(x.status) != 0) ? 1 : 0;
  __syncthreads();
  if (s_xch_fail64) return false;  // the run is lost already: neither publish nor wait
  for (int i = threadIdx.x; i < n * W; i += blockDim.x) {
    const int r = i / W, t = i - r * W;
    reinterpret_cast<unsigned long long*>(x.peer[r])[(static_cast<size_t>(par) * n + me) * W + t] = __ldcg(in + t);
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
        s_xch_fail64 = 1;
        break;
      }
    }
  }
  __syncthreads();
  if (s_xch_fail64) {
    __threadfence();  // the flag is visible to the rest of the grid before anyone acts on it
    return false;
  }
  for (int t = threadIdx.x; t < W; t += blockDim.x) {
    unsigned long long sum = 0;
    for (int r = 0; r < n; ++r) sum += __ldcg(reinterpret_cast<const unsigned long long*>(x.peer[me]) + (static_cast<size_t>(par) * n + r) * W + t);
    out[t] = sum;
  }
  __syncthreads();
  return true;
}

template <bool TRACE, bool INIT>
__global__ void __maxnreg__(HTM_GIBBS_MAXREG)
    gibbs_f32_kernel(const GibbsParams<float> p, const GibbsDecide d, const F32State st, const int iter_first,
                     const int iter_last, const int rec_origin, const int rec_cap, htm_step_trace* trace_base,
                     htm_swap_trace* swap_base, unsigned long long* accs /* [3][cur, prop][J][2 limbs], zero at launch */,
                     const int n_oct, unsigned long long* totals /* [cur, prop][J][2 limbs]: sums over all shards */,
                     const uint64_t seed, const int n_stages) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cg::grid_group grid = cg::this_grid();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  const int S = p.S, J = p.J, E = p.E, n_pairs = S / 2;
  const int nc = n_warps * kQuad;
  // the chain quads are spread EVENLY over the CTA rows (e.g. 25 quads over 4 rows: 7, 6, 6, 6 -- not 7, 7, 7, 4), so
  // that the SMs, which hold CTAs of different rows, carry the same number of active warps
  const int n_quads = (J + kQuad - 1) / kQuad;
  const int q_begin = static_cast<int>(static_cast<long>(n_quads) * blockIdx.y / gridDim.y);
  const int q_end = static_cast<int>(static_cast<long>(n_quads) * (blockIdx.y + 1) / gridDim.y);
  const int c_base = q_begin * kQuad, c_end = min(J, q_end * kQuad);
  const bool writer = blockIdx.x == 0 && blockIdx.y == 0;
  const F32Sm m = carve_f32_sm(smem_raw, S, p.xrow, nc, n_stages);
  const ChainSm cs = carve_chain_sm_small(smem_raw + ((f32_sweep_smem(S, nc, n_stages) + 15) & ~static_cast<size_t>(15)), J,
                                          d.g_tc, d.g_ac);
  const int o_begin = static_cast<int>(static_cast<long>(n_oct) * blockIdx.x / gridDim.x);
  const int o_end = static_cast<int>(static_cast<long>(n_oct) * (blockIdx.x + 1) / gridDim.x);
  const int n_my = o_end - o_begin;  // >= 1: the launcher never starts more CTAs per chain group than octets
  const int n_it = INIT ? 1 : iter_last - iter_first + 1;
  const long n_run = static_cast<long>(n_my) * n_it;  // octet visits of this CTA
  const uint32_t row_bytes = static_cast<uint32_t>(p.xrow * sizeof(float4));
  // warps whose four chain slots all lie past chain J-1 (the last chain group of a CTA row) stay out of the ring:
  // they would only spin on its barriers
  const int n_active = q_end - q_begin;  // <= n_warps
  if (threadIdx.x == 0) {
    for (int b = 0; b < n_stages; ++b) {
      mbar_init(m.full + b, 1);
      mbar_init(m.empty + b, n_active);
    }
    fence_mbar_init();
    fence_proxy_async();
  }
  __syncthreads();
  // The ring indices advance by increments: visit t uses stage t mod n_stages with phase (t / n_stages) & 1 and octet
  // o_begin + t mod n_my, but t is a 64-bit count and the divisors are run-time values -- formed by division, these
  // cost every warp two 64-bit divide subroutines per visit (a quarter of the instructions outside the station loop).
  int iss_i = 0;  // producer: (next visit to issue) mod n_my
  auto issue = [&](const int buf) {  // the next visit of this CTA -> ring stage buf
    const int o = o_begin + iss_i;
    if (++iss_i == n_my) iss_i = 0;
    const int n_ev = min(kOct, E - o * kOct);
    if (lane == 0) mbar_expect_tx(m.full + buf, row_bytes * n_ev);
    __syncwarp();
    if (lane < n_ev)
      tma_load_1d(m.rows + (buf * kOct + lane) * m.row, p.obsx + static_cast<size_t>(o * kOct + lane) * p.xrow, row_bytes,
                  m.full + buf);
  };
  if (warp == 0) {
    for (int b = 0; b < n_stages; ++b)
      if (n_run > b) issue(b);
  }
  chain_load(d, cs, /*terms=*/false);
  __syncthreads();
  f32_stage_chain_terms(m, cs, nc, c_base, c_end, S);

  const int es = lane >> 2, chs = lane & (kQuad - 1);
  const bool warp_ok = warp < n_active;
  const bool c_ok = warp_ok && c_base + warp * kQuad + chs < c_end;
  const int lc = c_ok ? warp * kQuad + chs : warp * kQuad;
  const int c = warp_ok ? c_base + lc : 0;
  const size_t per_it = static_cast<size_t>(E + 1) * J;
  // this lane's rows of the state arrays (a (chain, event) is always visited by the same thread)
  const size_t cE = static_cast<size_t>(c) * E;
  float4* const gH = st.H + cE;
  float4* const gM = st.M + cE;
  float4* const gQ = st.Q + cE;
  float4* const gP = st.P + cE;
  float* const gLp = st.Lp + cE;
  float4* const pf0 = m.pf + static_cast<size_t>(threadIdx.x) * 4;                     // stage 0
  float4* const pf1 = pf0 + static_cast<size_t>(blockDim.x) * 4;                       // stage 1
  uint32_t cnt_p[3] = {0, 0, 0}, cnt_a[3] = {0, 0, 0};
  long t_run = 0;
  int buf = 0;       // t_run mod n_stages
  uint32_t ph = 0;   // (t_run / n_stages) & 1

  for (int it = iter_first; it < iter_first + n_it; ++it) {
    // three accumulator sets in rotation: this iteration's, the next one's (already zero) and the one the writer
    // zeroes after this iteration's barrier (every CTA has read it before arriving there)
    unsigned long long* const acc_it = accs + static_cast<size_t>(it % 3) * 4 * J;
    const bool rec = !INIT && p.n_interval > 1 && (it % p.n_interval) == 1;
    int rec_slot = rec ? (it - 1) / p.n_interval - rec_origin : -1;
    if (rec_slot >= rec_cap) rec_slot = -1;
    htm_step_trace* trace_it = trace_base ? trace_base + static_cast<size_t>(it - iter_first) * per_it : nullptr;
    const bool a_prev = !INIT && cs.aprev[c] != 0;
    // this lane's state of octet visit i -> shared memory, asynchronously (one group per visit): the L2 latency
    // of the state hides behind the previous visit's arithmetic
    auto prefetch = [&](int i) {
      if (INIT || !kPrefetch) return;
      const int ee = min((o_begin + i) * kOct + es, E - 1);
      float4* dst = (i & 1) ? pf1 : pf0;
      cp_async16(dst, gH + ee);
      cp_async16(dst + 1, gM + ee);
      cp_async16(dst + 2, (a_prev ? gP : gQ) + ee);
      if (a_prev) cp_async4(dst + 3, gLp + ee);
      cp_async_commit();
    };
    HTM_PHASE(0);
    if (warp_ok) prefetch(0);
    // coefficients of the chains' pending proposals (decide_core ended with a block barrier; the previous
    // iteration's reads of pq are over)
    if (c_base + threadIdx.x < c_end && !INIT) f32_stage_proposal(m, cs, threadIdx.x, c_base + threadIdx.x, S);
    __syncthreads();
    const float T = static_cast<float>(cs.T[c]);
    const float iT = 1.f / T;
    const bool cold = gibbs_is_cold<float>(cs.T[c]);
    const Glob<float> g = make_glob<float>(static_cast<float>(cs.vs[c]), static_cast<float>(cs.qs[c]));
    PropF32 pr;
    pr.which = INIT ? 0 : cs.which[c];
    pr.idx = cs.idx[c];
    pr.qa = INIT ? 0.f : m.pq[3 * lc];
    pr.qb = INIT ? 0.f : m.pq[3 * lc + 1];
    pr.dlt = INIT ? 0.f : m.pq[3 * lc + 2];
    const bool any_station = __any_sync(0xffffffffu, pr.which == 2 || pr.which == 4);
    const float4* cp = m.cp + lc * m.cps;
    const float4 c0 = warp_ok ? m.c0[lc] : make_float4(0.f, 0.f, 0.f, 0.f);
    const int rec_chain_slot = (rec_slot >= 0 && p.hypo_rec) ? cs.slot[c] : -1;
    float4* const rec_row =
        rec_chain_slot >= 0 ? p.hypo_rec + (static_cast<size_t>(rec_slot) * p.n_cool_total + rec_chain_slot) * E : nullptr;
    double s_cur = 0.0, s_prop = 0.0;
    HTM_PHASE(1);
#ifdef HTM_GIBBS_PHASE_TRACE
    if (!INIT && it - iter_first == 10 && threadIdx.x == 0 && blockIdx.y * gridDim.x + blockIdx.x < 4096)
      g_cta_done_ns[2 * (blockIdx.y * gridDim.x + blockIdx.x)] = global_timer_ns();
#endif
    for (int i = 0; i < (warp_ok ? n_my : 0); ++i, ++t_run) {
      const int o = o_begin + i;
      if (warp_ok) {
        if (i + 1 < n_my) {
          prefetch(i + 1);
          cp_async_wait<1>();
        } else {
          cp_async_wait<0>();
        }
      }
      mbar_wait(m.full + buf, ph);
      if (warp_ok) {
        const int n_ev = min(kOct, E - o * kOct);
        const int e = o * kOct + es;
        const bool ev_ok = e < E;
        const int ee = ev_ok ? e : E - 1;  // idle lanes clone the last event and never write
        const float4* row = m.rows + (buf * kOct + (ev_ok ? es : n_ev - 1)) * m.row;
        const float4 evc = row[2];
        const float4 mu = row[3];
        const uint32_t gid = (static_cast<uint32_t>(ee) + p.event_offset) * p.J_total + p.chain_offset + static_cast<uint32_t>(c);
        float4 H, Mm, Q;
        bool acc = false, dirty = false;
        int icmp = 0;
        if (INIT) {
          const u32x4 a = philox4x32_10(seed, 0u, gid, PHX_INIT, 0u);
          const u32x4 b = philox4x32_10(seed, 1u, gid, PHX_INIT, 0u);
          H.x = mu.x + M<float>::gauss(a.v[0], a.v[1]) * p.width_xy;
          H.y = mu.y + M<float>::gauss(a.v[2], a.v[3]) * p.width_xy;
          H.z = p.prior_z + M<float>::sqrt(-2.f * M<float>::log(M<float>::u_oo(b.v[0]))) * p.width_z;
          const Moments o1 = eval_moments_f32(row, n_pairs, H.x - mu.x, H.y - mu.y, H.z, g, cp, c0.x, c0.y, evc);
          H.w = o1.L;
          Mm = make_float4(o1.A1t, o1.A3t, o1.A1a, o1.A3a);
          Q = make_float4(o1.S1t, o1.S1a, o1.A2t, o1.A2a);
          dirty = acc = true;
        } else {
          if (kPrefetch) {
            const float4* src = (i & 1) ? pf1 : pf0;
            H = src[0];
            Mm = src[1];
            Q = src[2];
            if (a_prev) H.w = reinterpret_cast<const float*>(src + 3)[0];
          } else {
            H = __ldcg(gH + ee);
            Mm = __ldcg(gM + ee);
            Q = __ldcg((a_prev ? gP : gQ) + ee);
            if (a_prev) H.w = __ldcg(gLp + ee);
          }
          dirty = a_prev;  // lazy commit of the last shared-parameter acceptance: exactly what was judged
          // model_perturb for one coordinate (src/cls_model.f90:162-190); icmp 0 -> z, 1 -> y, 2 -> x
          const u32x4 w = philox4x32_10(p.rk, static_cast<uint32_t>(it), gid, PHX_STEP, 0u);
          icmp = static_cast<int>(below(w.v[0], 3u));
          const float gs = M<float>::gauss(w.v[1], w.v[2]);
          const bool isz = icmp == 0;
          const float x_old = isz ? H.z : (icmp == 1 ? H.y : H.x);
          const float mu1 = isz ? p.prior_z : (icmp == 1 ? mu.y : mu.x);
          const float sigma = isz ? p.width_z : p.width_xy;
          const float x_new = x_old + gs * (isz ? p.step_z : p.step_xy);
          const float dn = x_new - mu1, dl = x_old - mu1;
          float lpr = -(dn * dn - dl * dl) / (2.f * sigma * sigma);
          bool ok = true;
          if (isz) {
            if (x_new <= mu1)
              ok = false;
            else
              lpr = lpr + M<float>::log(dn) - M<float>::log(dl);
          }
          const float nx = icmp == 2 ? x_new : H.x, ny = icmp == 1 ? x_new : H.y, nz = isz ? x_new : H.z;
          const Moments o1 = eval_moments_f32(row, n_pairs, nx - mu.x, ny - mu.y, nz, g, cp, c0.x, c0.y, evc);
          // mcmc_judge_model (src/cls_mcmc.f90:193-203)
          const float ratio = (o1.L - H.w) * iT + lpr;
          const float ru = M<float>::u_co(w.v[3]);
          acc = ok && (ru > 0.f) && (M<float>::log(ru) <= ratio);
          if (acc) {
            H = make_float4(nx, ny, nz, o1.L);
            Mm = make_float4(o1.A1t, o1.A3t, o1.A1a, o1.A3a);
            Q = make_float4(o1.S1t, o1.S1a, o1.A2t, o1.A2a);
            dirty = true;
          }
          if (TRACE) {
            if (ev_ok && c_ok && trace_it) {
              htm_step_trace t;
              t.proposal_type = 5 + icmp;
              t.index = 3 * (e + 1) - icmp;
              t.prior_ok = ok ? 1 : 0;
              t.accepted = acc ? 1 : 0;
              t.log_likelihood = static_cast<double>(H.w);
              trace_it[static_cast<size_t>(e) * J + c] = t;
            }
          }
        }
        // the chain's pending shared-parameter proposal, for this event: O(1)
        const DeltaSums ds = pending_delta(pr, any_station, row, H.x - mu.x, H.y - mu.y, H.z, g, cp, c0.x, c0.y, Mm, Q);
        const float dchi = (ds.dS2t - ds.dS1t * fmaf(2.f, Q.x, ds.dS1t) * evc.y) + (ds.dS2a - ds.dS1a * fmaf(2.f, Q.y, ds.dS1a) * evc.z);
        const float dL = -0.5f * dchi;
        if (ev_ok && c_ok) {
          if (dirty) {
            gH[e] = H;
            gQ[e] = Q;
          }
          if (acc) gM[e] = Mm;
          gP[e] = make_float4(Q.x + ds.dS1t, Q.y + ds.dS1a, Q.z + ds.dA2t, Q.w + ds.dA2a);
          gLp[e] = H.w + dL;
          if (rec_row) rec_row[e] = H;
          s_cur += static_cast<double>(H.w);
          s_prop += static_cast<double>(H.w) + static_cast<double>(dL);
          if (cold && !INIT) {
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
      if (warp == 0 && t_run + n_stages < n_run) {
        if (lane == 0) mbar_wait(m.empty + buf, ph);
        __syncwarp();
        issue(buf);  // visit t_run + n_stages lands in the stage this visit has just released
      }
      if (++buf == n_stages) {
        buf = 0;
        ph ^= 1u;
      }
    }
    if (warp_ok) {
#pragma unroll
      for (int off = 16; off >= kQuad; off >>= 1) {
        s_cur += __shfl_xor_sync(0xffffffffu, s_cur, off);
        s_prop += __shfl_xor_sync(0xffffffffu, s_prop, off);
      }
      if (es == 0 && c_ok) {
        publish_sum(acc_it + static_cast<size_t>(c) * 2, s_cur);
        publish_sum(acc_it + (static_cast<size_t>(J) + c) * 2, s_prop);
      }
    }
    HTM_PHASE(2);
#ifdef HTM_GIBBS_PHASE_TRACE
    if (!INIT && it - iter_first == 10) {
      __syncthreads();
      const unsigned int cta = blockIdx.y * gridDim.x + blockIdx.x;
      if (threadIdx.x == 0 && cta < 4096) {
        g_cta_done_ns[2 * cta + 1] = global_timer_ns();
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_cta_info[4 * cta] = smid;
        g_cta_info[4 * cta + 1] = blockIdx.y;
        g_cta_info[4 * cta + 2] = static_cast<unsigned int>(n_my);
        g_cta_info[4 * cta + 3] = static_cast<unsigned int>(n_active);
      }
    }
#endif
    // one CTA (tiny problems, e.g. BASELINE configs[0]): a block barrier orders the partial sums as well
    if (gridDim.x * gridDim.y > 1) {
      grid.sync();
    } else {
      __threadfence();
      __syncthreads();
    }
    HTM_PHASE(3);
    if (writer) {
      unsigned long long* const z = accs + static_cast<size_t>((it + 2) % 3) * 4 * J;
      for (int t = threadIdx.x; t < 4 * J; t += blockDim.x) z[t] = 0ull;
    }
    // cs.tot[0..J) = sum_e L_e, [J..2J) = the same under the pending proposal; identical in every CTA
    auto load_sums = [&](const unsigned long long* src) {
      for (int t = threadIdx.x; t < 2 * J; t += blockDim.x) cs.tot[t] = limbs_to_double(__ldcg(src + 2 * t), __ldcg(src + 2 * t + 1));
      __syncthreads();
    };
    if (INIT) {  // g_L[c] = sum_e L_e
      if (writer) {
        load_sums(acc_it);
        for (int t = threadIdx.x; t < J; t += blockDim.x) d.g_L[t] = cs.tot[t];
      }
      break;
    }
    htm_step_trace* trace_g = trace_it ? trace_it + static_cast<size_t>(E) * J : nullptr;
    htm_swap_trace* swap_it = swap_base ? swap_base + (it - iter_first) : nullptr;
    // the proposal being judged, of the chain this thread looks after (decide_core replaces it by the next one)
    const bool mine = c_base + threadIdx.x < c_end;
    const int j_which = mine ? cs.which[c_base + threadIdx.x] : 0, j_idx = mine ? cs.idx[c_base + threadIdx.x] : 0;
    const double j_xnew = mine ? cs.xnew[c_base + threadIdx.x] : 0.0;
    if (d.xch.n > 1) {
      // event shards: CTA (0,0) exchanges this shard's sums with the other GPUs through peer memory and hands the
      // totals over all events to every CTA of its grid
      if (writer) {
        peer_allreduce_u64(d.xch, d.xch.epoch + static_cast<uint32_t>(it - iter_first), acc_it, totals, 4 * J);
        __threadfence();
      }
      HTM_PHASE(4);
      grid.sync();
      HTM_PHASE(5);
      // a peer that never answered ends the run here, on every CTA alike (only this shard's writer sets the
      // flag, before the barrier): no decision is taken from partial sums; the host reports HTM_ERR_CUDA
      if (*reinterpret_cast<volatile int*>(d.xch.status) != 0) break;
      load_sums(totals);
    } else {
      load_sums(acc_it);
    }
    decide_core(d, cs, it, it + 1, nullptr, nullptr, rec_slot, trace_g, swap_it, writer, true, true);
    // an accepted station-term proposal changes one entry of the chain's terms in shared memory
    if (mine && cs.aprev[c_base + threadIdx.x] && (j_which == 2 || j_which == 4))
      f32_update_chain_term(m, threadIdx.x, j_which, j_idx, j_xnew);
    HTM_PHASE(6);
  }
  if (INIT) return;
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
  if (writer) chain_store(d, cs, /*terms=*/false);
}

// cold slots + the proposal of the first iteration (a pure function of state and iteration number): one CTA
__global__ void __launch_bounds__(256) gibbs_f32_prepare_kernel(const GibbsDecide d) {
  extern __shared__ __align__(16) unsigned char s_prep_dyn[];
  gibbs_decide_small(d, s_prep_dyn);
}
static cudaError_t launch_gibbs_prepare(const GibbsLaunch& a, cudaStream_t stream) {
  const size_t sm = chain_sm_small_bytes(a.J);
  if (sm > 200 * 1024) return cudaErrorInvalidConfiguration;
  cudaError_t err = cudaFuncSetAttribute(gibbs_f32_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm));
  if (err != cudaSuccess) return err;
  GibbsDecide d = make_decide(a);
  d.it = 0;
  d.it_next = a.iter_first;
  gibbs_f32_prepare_kernel<<<1, 256, sm, stream>>>(d);
  return cudaGetLastError();
}

// ---- host launcher -----------------------------------------------------------------------------------------
struct F32Shape {
  int n_warps = kCW, gy = 1, n_oct = 1, n_stages = 2;
  long gx = 1;
  size_t smem = 0;
};
template <bool TRACE, bool INIT>
static cudaError_t f32_shape(const GibbsLaunch& a, F32Shape* s) {
  const int quads = (a.J + kQuad - 1) / kQuad;
  // CTA rows (chain groups).  Every row streams the whole event table, so few rows are cheap in traffic; but the
  // warps a row's CTAs carry decide how many CTAs fit an SM (registers, shared memory), and 25 chain quads in 4 rows
  // of 7 + 6 + 6 + 6 warps leave an SM with 12-13 resident warps where 5 rows of 5 warps give it 15.  Take the fewest
  // rows unless more rows raise the resident warps per SM by more than 8 % (sample shape, 100 joint chains:
  // 111 -> 105 us per iteration at 10 000 events, 1029 -> 930 us at 100 000; profiles/r2bi_gibbs_rows.txt).
  s->gy = (quads + kCW - 1) / kCW;
  {
    double best = 0.0;
    int last_gy = 0;
    for (int target = kCW; target >= 4; --target) {
      const int gy = (quads + target - 1) / target;
      if (gy == last_gy) continue;
      last_gy = gy;
      const int nw = (quads + gy - 1) / gy;
      const size_t smem = ((f32_sweep_smem(a.S, nw * kQuad, 2) + 15) & ~static_cast<size_t>(15)) + chain_sm_small_bytes(a.J);
      if (smem > 200 * 1024) continue;
      if (cudaFuncSetAttribute(gibbs_f32_kernel<TRACE, INIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) continue;
      int occ = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gibbs_f32_kernel<TRACE, INIT>, nw * 32, smem) != cudaSuccess) continue;
      const double warps_per_sm = static_cast<double>(occ) * quads / gy;
      if (warps_per_sm > best * 1.08) {
        best = warps_per_sm;
        s->gy = gy;
      }
    }
  }
  if (const char* v = std::getenv("HTM_GIBBS_ROWS")) s->gy = std::max((quads + kCW - 1) / kCW, std::min(quads, std::atoi(v)));  // tuning
  s->n_warps = (quads + s->gy - 1) / s->gy;
  // ring depth: as deep as possible (fast warps may then run ahead of the slowest one, and the row copies have
  // several visits to land) WITHOUT lowering the number of resident CTAs per SM that two stages allow
  int per_sm = 0, dev = 0, n_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  int per_sm2 = 0, max_stages = kMaxStages;
  if (const char* v = std::getenv("HTM_GIBBS_RING")) max_stages = std::max(2, std::min(kMaxStages, std::atoi(v)));  // tuning
  cudaError_t err = cudaSuccess;
  for (int ns : {2, max_stages, 3}) {
    if (ns < 2 || ns > max_stages || (ns == 2 && per_sm2 > 0)) continue;
    const size_t smem = ((f32_sweep_smem(a.S, s->n_warps * kQuad, ns) + 15) & ~static_cast<size_t>(15)) + chain_sm_small_bytes(a.J);
    if (smem > 200 * 1024) {
      if (ns == 2) return cudaErrorInvalidConfiguration;
      continue;
    }
    err = cudaFuncSetAttribute(gibbs_f32_kernel<TRACE, INIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (err != cudaSuccess) return err;
    int occ = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gibbs_f32_kernel<TRACE, INIT>, s->n_warps * 32, smem);
    if (err != cudaSuccess) return err;
    if (ns == 2) {
      per_sm2 = occ;
      s->n_stages = 2;
      s->smem = smem;
      per_sm = occ;
    } else if (occ >= per_sm2) {
      s->n_stages = ns;
      s->smem = smem;
      per_sm = occ;
      break;
    }
  }
  err = cudaFuncSetAttribute(gibbs_f32_kernel<TRACE, INIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(s->smem));
  if (err != cudaSuccess) return err;
  s->n_oct = (a.E + kOct - 1) / kOct;
  s->gx = static_cast<long>(per_sm) * n_sm / s->gy;
  if (s->gx > s->n_oct) s->gx = s->n_oct;
  if (s->gx < 1) return cudaErrorCooperativeLaunchTooLarge;  // more chain groups than resident CTAs
  return cudaSuccess;
}

template <bool TRACE, bool INIT>
static cudaError_t launch_f32(const GibbsLaunch& a, cudaStream_t stream) {
  F32Shape s;
  cudaError_t err = f32_shape<TRACE, INIT>(a, &s);
  if (err != cudaSuccess) return err;
  GibbsParams<float> pp = make_gibbs_params<float>(a);
  GibbsDecide dp = make_decide(a);
  pp.n_tiles = dp.n_tiles = static_cast<int>(s.gx);
  if (a.out_partials) *a.out_partials = static_cast<int>(s.gx);
  dp.xch.epoch = a.xch_epoch0;  // exchange number of iter_first; the kernel counts on from there
  F32State st;
  st.H = static_cast<float4*>(a.sH);
  st.M = static_cast<float4*>(a.sM);
  st.Q = static_cast<float4*>(a.sQ);
  st.P = static_cast<float4*>(a.sP);
  st.Lp = static_cast<float*>(a.hLp);
  int iter_first = a.iter_first, iter_last = a.iter_last, rec_origin = a.rec_origin, rec_cap = a.rec_cap, n_oct = s.n_oct;
  htm_step_trace* tr = a.trace;
  htm_swap_trace* sw = a.swaps;
  unsigned long long* accs = reinterpret_cast<unsigned long long*>(a.part_cur);
  unsigned long long* totals = reinterpret_cast<unsigned long long*>(a.totals);
  err = cudaMemsetAsync(accs, 0, static_cast<size_t>(12) * a.J * sizeof(unsigned long long), stream);
  if (err != cudaSuccess) return err;
  uint64_t seed = a.seed;
  int n_stages = s.n_stages;
  void* args[] = {&pp, &dp, &st, &iter_first, &iter_last, &rec_origin, &rec_cap, &tr, &sw, &accs, &n_oct, &totals, &seed, &n_stages};
  return cudaLaunchCooperativeKernel(reinterpret_cast<void*>(gibbs_f32_kernel<TRACE, INIT>),
                                     dim3(static_cast<unsigned>(s.gx), s.gy), dim3(s.n_warps * 32), args, s.smem, stream);
}

double gibbs_f32_sum_to_double(unsigned long long l0, unsigned long long l1) { return limbs_to_double(l0, l1); }

// float32 mode C run: prepare (cold slots + first proposal) + ONE cooperative launch for all iterations
cudaError_t launch_gibbs_f32(const GibbsLaunch& a, cudaStream_t stream, int* n_launches) {
  if (a.comm && a.xch.n <= 1) return cudaErrorNotSupported;  // float32 event shards exchange through peer memory only
  cudaError_t err = launch_gibbs_prepare(a, stream);
  if (err != cudaSuccess) return err;
  err = (a.trace || a.swaps) ? launch_f32<true, false>(a, stream) : launch_f32<false, false>(a, stream);
  if (err != cudaSuccess) return err;
  if (n_launches) *n_launches = 2;
  return cudaGetLastError();
}

#ifdef HTM_GIBBS_PHASE_TRACE
}  // namespace htm
extern "C" int32_t htm_debug_cta_info(unsigned int* out, int32_t n_cta) {
  return cudaMemcpyFromSymbol(out, htm::g_cta_info, static_cast<size_t>(n_cta) * 4 * sizeof(unsigned int)) == cudaSuccess ? 0 : 3;
}

extern "C" int32_t htm_debug_cta_trace(unsigned long long* out, int32_t n_cta) {
  if (n_cta > 4096) n_cta = 4096;
  return cudaMemcpyFromSymbol(out, htm::g_cta_done_ns, static_cast<size_t>(n_cta) * 2 * sizeof(unsigned long long)) == cudaSuccess ? 0 : 3;
}
extern "C" int32_t htm_debug_phase_trace(unsigned long long* out, int32_t n_iter) {
  if (n_iter > 4096) n_iter = 4096;
  return cudaMemcpyFromSymbol(out, htm::g_phase_ns, static_cast<size_t>(n_iter) * 8 * sizeof(unsigned long long)) == cudaSuccess ? 0 : 3;
}
namespace htm {
#endif

// hypocentres + state of every (chain, event) at the initial shared parameters, and g_L
cudaError_t launch_gibbs_f32_init(const GibbsLaunch& a, cudaStream_t stream) {
  GibbsLaunch b = a;
  b.iter_first = b.iter_last = 0;
  b.trace = nullptr;
  b.swaps = nullptr;
  b.xch = PeerExchange();
  return launch_f32<false, true>(b, stream);
}

}  // namespace htm
