// Mode A: the reference-exact joint chain, driven by the host's mod_random draws.
//
// One chain = ALL events + vs + qs + t_corr(S) + a_corr(S); one scalar is perturbed per
// iteration (src/cls_mcmc.f90:115-172); chains of a virtual rank share ONE draw stream and
// are stepped in order; one swap attempt per iteration over all n_procs*n_chains chains
// (src/cls_parallel.f90:100-240).  The number of draws a step consumes depends on the data
// (3-6), so a rank is inherently sequential: this is a VALIDATION mode (float64, one CTA,
// one warp per virtual rank, stations across lanes) and is not optimised.
//
// Draw schedule per (iteration, chain), src/cls_mcmc.f90:134-165 + src/cls_model.f90:172:
//   a_select (rand_u) ; [station id | event id, component] (rand_u) ; v1, v2 (rand_u2) ;
//   r (rand_u, only when prior_ok).  Then rank 0: >= 2 draws in select_pair; rank1: 1 draw.
#include "htm_forward.cuh"
#include "htm_kernels.hpp"

namespace htm {

// (dble(w) + 2^31) / 2^32 and (dble(w) + 2^31 + 0.5) / 2^32, src/mod_random.f90:72,90
__device__ __forceinline__ double to_u(int32_t w) { return (static_cast<double>(w) + 2147483648.0) / 4294967296.0; }
__device__ __forceinline__ double to_u2(int32_t w) {
  return (static_cast<double>(w) + 2147483648.0 + 0.5) / 4294967296.0;
}

// log-likelihood contribution of one event for chain c, with an optional proposed override
__device__ __forceinline__ double rp_event_L(const ReplayLaunch& a, int c, int e, double px, double py, double pz,
                                             const Glob<double>& g, int ov_which, int ov_idx, double ov_val) {
  const double4* sta4 = static_cast<const double4*>(a.tab.sta4);
  const double4* obs4 = static_cast<const double4*>(a.tab.obs4_raw) + static_cast<size_t>(e) * a.S;
  const double4 evc = static_cast<const double4*>(a.tab.evc4)[e];
  return warp_event_loglik<double, double>(sta4, obs4, evc, a.S, px, py, pz, g, a.tc + static_cast<size_t>(c) * a.S,
                                           a.ac + static_cast<size_t>(c) * a.S, ov_which, ov_idx, ov_val);
}

// forward%calc_log_likelihood (src/cls_forward.f90:268-303) for chain c with the proposal applied
__device__ double rp_full_L(const ReplayLaunch& a, int c, const Glob<double>& g, int which, int idx, double x_new) {
  double L = 0.0;
  const double* h = a.hypo + static_cast<size_t>(c) * 3 * a.E;
  for (int e = 0; e < a.E; ++e) {
    double px = h[3 * e], py = h[3 * e + 1], pz = h[3 * e + 2];
    if (which == 5 && idx / 3 == e) {
      const int comp = idx - 3 * e;
      if (comp == 0) px = x_new;
      if (comp == 1) py = x_new;
      if (comp == 2) pz = x_new;
    }
    L += rp_event_L(a, c, e, px, py, pz, g, which, (which == 2 || which == 4) ? idx : -1, x_new);
  }
  return L;
}

__global__ void __launch_bounds__(256) replay_kernel(const ReplayLaunch a) {
  extern __shared__ long long s_cursor[];  // [R]
  __shared__ int s_abort;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int R = a.R, K = a.K, E = a.E, S = a.S;
  for (int r = threadIdx.x; r < R; r += blockDim.x) s_cursor[r] = a.cursor[r];
  if (threadIdx.x == 0) s_abort = 0;
  __syncthreads();
  const double pi2 = 2.0 * kPi;
  const double t1 = a.p_vs, t2 = a.p_vs + a.p_t_corr, t3 = a.p_vs + a.p_t_corr + a.p_qs,
               t4 = a.p_vs + a.p_t_corr + a.p_qs + a.p_a_corr;  // src/cls_mcmc.f90:139-153

  for (int it = a.iter_first; it <= a.iter_last; ++it) {
    for (int r = warp; r < R; r += nw) {
      const int32_t* dr = a.draws + a.draw_off[r];
      const long long n_avail = a.draw_off[r + 1] - a.draw_off[r];
      for (int j = 0; j < K; ++j) {
        const int c = r * K + j;
        long long cur = s_cursor[r];
        bool starved = false;  // warp-uniform: the supplied stream ran out
        auto next = [&]() -> int32_t {
          if (cur >= n_avail) {
            starved = true;
            return 0;
          }
          return dr[cur++];
        };
        // ---- mcmc_propose_model ----
        const double a_select = to_u(next());
        int which, idx, type, evt = -999;
        if (a_select < t1) {
          which = 1; idx = 0; type = 1;
        } else if (a_select < t2) {
          which = 2; idx = static_cast<int>(to_u(next()) * S); type = 2;
        } else if (a_select < t3) {
          which = 3; idx = 0; type = 3;
        } else if (a_select < t4) {
          which = 4; idx = static_cast<int>(to_u(next()) * S); type = 4;
        } else {
          const int id = static_cast<int>(to_u(next()) * E) + 1;
          const int icmp = static_cast<int>(to_u(next()) * 3);
          which = 5; idx = 3 * id - icmp - 1; type = 5 + icmp; evt = id;
        }
        // ---- model_perturb (src/cls_model.f90:162-190) ----
        double* slot;
        double mu, sigma, step;
        int prior_type = 0;
        if (which == 1) {
          slot = a.vs + c; mu = a.prior_vs; sigma = a.width_vs; step = a.step_vs;
        } else if (which == 2) {
          slot = a.tc + static_cast<size_t>(c) * S + idx; mu = a.prior_tc; sigma = a.width_tc; step = a.step_tc;
        } else if (which == 3) {
          slot = a.qs + c; mu = a.prior_qs; sigma = a.width_qs; step = a.step_qs;
        } else if (which == 4) {
          slot = a.ac + static_cast<size_t>(c) * S + idx; mu = a.prior_ac; sigma = a.width_ac; step = a.step_ac;
        } else {
          slot = a.hypo + static_cast<size_t>(c) * 3 * E + idx;
          const int e = idx / 3, comp = idx - 3 * e;
          if (comp == 2) {
            mu = a.prior_z; sigma = a.width_z; step = a.step_z; prior_type = 1;
          } else {
            mu = a.prior_xy[2 * e + comp]; sigma = a.width_xy; step = a.step_xy;
          }
        }
        const double v1 = to_u2(next());
        const double v2 = to_u2(next());
        if (starved) {
          if (lane == 0) s_abort = 1;
          break;
        }
        const double gs = ::sqrt(-2.0 * ::log(v1)) * ::cos(pi2 * v2);  // rand_g
        const double x_old = *slot;
        const double x_new = __dadd_rn(x_old, __dmul_rn(gs, step));
        const double dn = x_new - mu, dl = x_old - mu;
        double lpr = -(__dmul_rn(dn, dn) - __dmul_rn(dl, dl)) / (2.0 * sigma * sigma);
        bool prior_ok = true;
        if (prior_type == 1) {
          if (x_new <= mu) {
            prior_ok = false;
          } else {
            lpr = lpr + ::log(dn) - ::log(dl);
          }
        }
        // ---- forward (src/hypo_tremor_mcmc.f90:245-259) ----
        const double L_cur = a.L[c];
        double L_new = 0.0;
        if (prior_ok) {
          const double vsv = which == 1 ? x_new : a.vs[c];
          const double qsv = which == 3 ? x_new : a.qs[c];
          const Glob<double> g = make_glob<double>(vsv, qsv);
          if (evt > 0 && it > 1) {
            // partially_update_log_likelihood (src/cls_forward.f90:307-362): remove the old
            // event's terms, add the new ones
            const int e = evt - 1, comp = idx - 3 * e;
            const double* h = a.hypo + static_cast<size_t>(c) * 3 * E + 3 * e;
            const double ox = h[0], oy = h[1], oz = h[2];
            const double nx = comp == 0 ? x_new : ox, ny = comp == 1 ? x_new : oy, nz = comp == 2 ? x_new : oz;
            const double Le_old = rp_event_L(a, c, e, ox, oy, oz, g, 0, -1, 0.0);
            const double Le_new = rp_event_L(a, c, e, nx, ny, nz, g, 0, -1, 0.0);
            L_new = (L_cur - Le_old) + Le_new;
          } else {
            L_new = rp_full_L(a, c, g, which, idx, x_new);
          }
        }
        // ---- mcmc_judge_model (src/cls_mcmc.f90:176-226) ----
        const double temp = a.temp[c];
        const bool cold = temp < 1.0 + kEps64;
        bool acc = false;
        if (prior_ok) {
          double ratio = (L_new - L_cur) / temp;
          ratio = ratio + lpr;
          const double rr = to_u(next());
          if (rr >= kEps64) {
            if (::log(rr) <= ratio) acc = true;
          }
        }
        if (starved) {
          if (lane == 0) s_abort = 1;
          break;
        }
        __syncwarp();  // every lane has read the old state
        if (lane == 0) {
          unsigned long long* cnt = a.chain_counts + static_cast<size_t>(c) * 14;
          if (cold) cnt[type - 1] += 1;
          if (acc) {
            *slot = x_new;
            a.L[c] = L_new;
            if (cold) cnt[7 + type - 1] += 1;
          }
          if (a.trace) {
            htm_step_trace t;
            t.proposal_type = type;
            t.index = idx + 1;
            t.prior_ok = prior_ok ? 1 : 0;
            t.accepted = acc ? 1 : 0;
            t.log_likelihood = acc ? L_new : L_cur;
            a.trace[(static_cast<size_t>(it - a.iter_first) * R + r) * K + j] = t;
          }
          s_cursor[r] = cur;
        }
        __syncwarp();
      }
    }
    __syncthreads();
    // ---- parallel_swap_temperature (src/cls_parallel.f90:100-240), by one thread ----
    if (threadIdx.x == 0 && !s_abort && R * K >= 2) {
      const int32_t* d0 = a.draws + a.draw_off[0];
      const long long n0 = a.draw_off[1] - a.draw_off[0];
      long long c0 = s_cursor[0];
      bool ok = c0 + 2 <= n0;
      int i1 = 0, i2 = 0;
      if (ok) {
        i1 = static_cast<int>(to_u(d0[c0++]) * R * K);  // rand_u()*n_proc*n_chain, left to right
        for (;;) {
          if (c0 >= n0) {
            ok = false;
            break;
          }
          i2 = static_cast<int>(to_u(d0[c0++]) * R * K);
          if (i1 != i2) break;
        }
      }
      if (ok) {
        s_cursor[0] = c0;
        const int rank1 = i1 / K, rank2 = i2 / K, ch1 = i1 % K, ch2 = i2 % K;
        const int c1 = rank1 * K + ch1, c2 = rank2 * K + ch2;
        const int32_t* d1 = a.draws + a.draw_off[rank1];
        long long cu = s_cursor[rank1];
        if (cu + 1 > a.draw_off[rank1 + 1] - a.draw_off[rank1]) {
          ok = false;
        } else {
          const double rr = to_u(d1[cu++]);  // judge_swap draws on rank1's stream, :129,:163
          s_cursor[rank1] = cu;
          const double T1 = a.temp[c1], T2 = a.temp[c2], L1 = a.L[c1], L2 = a.L[c2];
          const double del_s = (L2 - L1) * (1.0 / T1 - 1.0 / T2);
          bool sacc = false;
          if (rr >= kEps64) {
            if (::log(rr) <= del_s) sacc = true;
          }
          if (sacc) {
            a.temp[c1] = T2;
            a.temp[c2] = T1;
          }
          if (a.swaps) {
            htm_swap_trace t;
            t.rank1 = rank1;
            t.chain1 = ch1 + 1;
            t.rank2 = rank2;
            t.chain2 = ch2 + 1;
            t.accepted = sacc ? 1 : 0;
            t.reserved = 0;
            a.swaps[it - a.iter_first] = t;
          }
        }
      }
      if (!ok) s_abort = 1;
    }
    __syncthreads();
    if (s_abort) break;
  }
  __syncthreads();
  for (int r = threadIdx.x; r < R; r += blockDim.x) a.cursor[r] = s_cursor[r];
  if (threadIdx.x == 0) *a.status = s_abort;
}

cudaError_t launch_replay(const ReplayLaunch& a, cudaStream_t stream) {
  int nw = a.R < 8 ? a.R : 8;  // ranks beyond 8 are looped over by the warps
  const size_t smem = static_cast<size_t>(a.R) * sizeof(long long);
  replay_kernel<<<1, nw * 32, smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace htm
