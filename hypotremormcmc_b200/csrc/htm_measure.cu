// Upstream stage hypo_tremor_measure, numerical core (second half): for every detected time
// window the relative arrival times and log-amplitudes of the station envelopes and their scatter
// (src/cls_measurer.f90:405-523, called from measure_lag_time :317-400).  One CTA per window.
//
// The reference correlates every station pair with FFTW (r2c, conjg(X_i) X_j, c2r) and takes maxloc.  Here the circular
// cross-correlation r_ij(k) = sum_m a_i(m) a_j((m + k) mod n) is evaluated directly in float64: at the reference's
// sizes (n = t_win / dt = 300 samples, 50 stations) that is 1.1e8 DFMA per window -- 6 us at the FP64 rate of a B200,
// no transform length restrictions, and no transform round-off in the quantity whose arg-max is taken.
//
// Layout: the window's tapered, normalised envelopes a[S][stride] in shared memory, each row followed by its own first
// 32 samples (so that reads up to 31 samples past any position below n never wrap) and padded to stride = 2 mod 16
// doubles (rows of consecutive stations start 16 B apart in the banks: a quarter-warp reading the same column of eight
// consecutive rows with LDS.128 is conflict-free).  A lane owns ONE station pair and a tile of 16 consecutive lags,
// whose 16 running sums and a 16-sample sliding window of a_j stay in registers: one step of m is 16 DFMA against one
// new sample of a_i and one of a_j (the loop is unrolled over 16 steps so that the window rotates through fixed
// register names) -- 0.125 shared-memory wavefronts per warp-DFMA, where the FP64 pipe needs <= 0.5.  The wrap position
// `base` = (m + k0) mod n is the same for the whole warp.  Pairs are taken 32 at a time in the reference's order, so
// the lanes of a warp mostly share station i (broadcast) and read 32 different rows j.
#include <cfloat>

#include "htm_kernels.hpp"

namespace htm {

namespace {

constexpr int kMeasureThreads = 512;
constexpr int kLagTile = 16;  // consecutive lags per lane = length of the a_j window = m steps per unrolled block
constexpr int kRowExt = 32;   // wrap extension of a row
#ifndef HTM_MEASURE_UNITS
#define HTM_MEASURE_UNITS 10
#endif
constexpr int kUnitsPerWarp = HTM_MEASURE_UNITS;  // units (32 pairs x a share of the lag tiles) per warp, at least

__host__ __device__ inline int measure_stride(int n) {
  int s = n + kRowExt;
  while ((s & 15) != 2) ++s;
  return s;
}

struct MeasureSm {
  double* rows;    // [S][stride]
  double* bestv;   // [n_split][P]; later rel [P]
  int* bestk;      // [n_split][P]
  int* lagk;       // [P]
  double* tS;      // [S] t, then amp
  double* sxx;     // [S]
  int* it;         // [S]
  uint16_t* pi;    // [P]
  uint16_t* pj;    // [P]
  int* flag;
};

__host__ __device__ inline size_t measure_smem(int S, int n, int n_split) {
  const size_t P = static_cast<size_t>(S) * (S - 1) / 2;
  size_t b = static_cast<size_t>(S) * measure_stride(n) * 8;  // rows
  b += n_split * P * 8;                                        // bestv
  b += n_split * P * 4 + P * 4;                                // bestk, lagk
  b = (b + 7) & ~static_cast<size_t>(7);
  b += 2 * static_cast<size_t>(S) * 8 + static_cast<size_t>(S) * 4 + 8;  // tS, sxx, it, flag[2]
  b += 2 * P * 2;                                                        // pi, pj
  return b + 16;
}

__device__ inline MeasureSm carve_measure(unsigned char* base, int S, int n, int n_split) {
  const size_t P = static_cast<size_t>(S) * (S - 1) / 2;
  MeasureSm m;
  unsigned char* q = base;
  m.rows = reinterpret_cast<double*>(q);
  q += static_cast<size_t>(S) * measure_stride(n) * 8;
  m.bestv = reinterpret_cast<double*>(q);
  q += n_split * P * 8;
  m.bestk = reinterpret_cast<int*>(q);
  q += n_split * P * 4;
  m.lagk = reinterpret_cast<int*>(q);
  q += P * 4;
  q = base + ((static_cast<size_t>(q - base) + 7) & ~static_cast<size_t>(7));
  m.tS = reinterpret_cast<double*>(q);
  q += static_cast<size_t>(S) * 8;
  m.sxx = reinterpret_cast<double*>(q);
  q += static_cast<size_t>(S) * 8;
  m.it = reinterpret_cast<int*>(q);
  q += static_cast<size_t>(S) * 4;
  m.flag = reinterpret_cast<int*>(q);
  q += 8;
  m.pi = reinterpret_cast<uint16_t*>(q);
  q += P * 2;
  m.pj = reinterpret_cast<uint16_t*>(q);
  return m;
}

__device__ __forceinline__ int pair_index(int S, int i, int j) { return i * (2 * S - i - 1) / 2 + (j - i - 1); }

// signed entry (i, j) of an antisymmetric table stored for i < j
__device__ __forceinline__ double antisym(const double* tab, int S, int i, int j) {
  if (i == j) return 0.0;
  return i < j ? tab[pair_index(S, i, j)] : -tab[pair_index(S, j, i)];
}

// src/cls_measurer.f90:499-503: 1-based maxloc position -> signed lag
__device__ __forceinline__ double lag_of(int k, int n, double dt) { return (k + 1 <= n / 2) ? k * dt : (k - n) * dt; }

// 16 steps of m for one pair and 16 lags: acc[t] += a_i(m + q) * a_j(base + q + t), q = 0..15.  W holds a_j(base + q ..
// base + q + 15) at step q, sample base + q + s' in slot (q + s') mod 16; the sample leaving the window is replaced by
// a_j(base + q + 16).  ri_m = a_i + m (16-byte aligned), rj_n = a_j + base + 16 (16-byte aligned unless kOdd).
// kLast: only the first n_valid steps exist (m + q < n); the others enter with a_i = 0.
template <bool kLast, bool kOdd>
__device__ __forceinline__ void corr_block(const double* __restrict__ ri_m, const double* __restrict__ rj_n,
                                           double (&W)[kLagTile], double (&acc)[kLagTile], const int n_valid) {
#pragma unroll
  for (int q2 = 0; q2 < kLagTile / 2; ++q2) {
    double2 ai = *reinterpret_cast<const double2*>(ri_m + 2 * q2);
    double2 nw;
    if (kOdd) {
      nw.x = rj_n[2 * q2];
      nw.y = rj_n[2 * q2 + 1];
    } else {
      nw = *reinterpret_cast<const double2*>(rj_n + 2 * q2);
    }
    if (kLast) {
      if (2 * q2 >= n_valid) ai.x = 0.0;
      if (2 * q2 + 1 >= n_valid) ai.y = 0.0;
    }
#pragma unroll
    for (int t = 0; t < kLagTile; ++t) acc[t] = fma(ai.x, W[(2 * q2 + t) & (kLagTile - 1)], acc[t]);
    W[(2 * q2) & (kLagTile - 1)] = nw.x;
#pragma unroll
    for (int t = 0; t < kLagTile; ++t) acc[t] = fma(ai.y, W[(2 * q2 + 1 + t) & (kLagTile - 1)], acc[t]);
    W[(2 * q2 + 1) & (kLagTile - 1)] = nw.y;
  }
}

}  // namespace

__global__ void __launch_bounds__(kMeasureThreads, 1) measure_kernel(const MeasureArgs a, const int n_split) {
  extern __shared__ __align__(16) unsigned char measure_smem_raw[];
  const int S = a.S, n = a.n, P = S * (S - 1) / 2, stride = measure_stride(n);
  const MeasureSm sm = carve_measure(measure_smem_raw, S, n, n_split);
  const int w = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  // src/cls_measurer.f90:331-333: window id -> first sample (0-based) of the merged envelopes
  const long j1 = static_cast<long>(a.win_id ? a.win_id[w] - 1 : w) * a.n_step;

  for (int i = threadIdx.x; i < S - 1; i += blockDim.x) {
    int p = pair_index(S, i, i + 1);
    for (int j = i + 1; j < S; ++j, ++p) {
      sm.pi[p] = static_cast<uint16_t>(i);
      sm.pj[p] = static_cast<uint16_t>(j);
    }
  }
  if (threadIdx.x == 0) sm.flag[0] = sm.flag[1] = 0;  // negative-product flag, next unit of the correlation phase

  // ---- the tapered, normalised windows -----------------------------------------------------------------------------
  // measure  (optimize_cc, src/cls_measurer.f90:478-486):          a_i = taper(x_i) / sum(x_i^2)
  // correlate (run_cross_corr, src/cls_correlator.f90:207-226):     a_i = (taper(x_i) - mean) / |taper(x_i) - mean|
  const int nleng = static_cast<int>(n * 0.05);
  const double pi_d = 3.14159265358979323846;
  for (int i = warp; i < S; i += n_warps) {
    const double* x = a.env + static_cast<size_t>(i) * a.n_total + j1;
    double* r = sm.rows + static_cast<size_t>(i) * stride;
    double l = 0.0, sum = 0.0;
    for (int m = lane; m < n; m += 32) {
      const int e = m < nleng ? m : (n - 1 - m < nleng ? n - 1 - m : -1);
      const double xv = x[m];
      double v = xv;
      if (e >= 0) v = v * (0.5 * (1.0 - cos(e * pi_d / nleng)));
      r[m] = v;
      l = fma(xv, xv, l);
      sum += v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      l += __shfl_xor_sync(0xffffffffu, l, o);
      sum += __shfl_xor_sync(0xffffffffu, sum, o);
    }
    __syncwarp();
    if (a.mode == 0) {
      for (int m = lane; m < n; m += 32) r[m] = r[m] / l;
    } else {
      const double mean = sum / n;
      double l2 = 0.0;
      for (int m = lane; m < n; m += 32) {
        const double v = r[m] - mean;
        l2 = fma(v, v, l2);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) l2 += __shfl_xor_sync(0xffffffffu, l2, o);
      const double len = sqrt(l2);
      // (a window of zeros: the reference leaves its work array untouched, :211-213; here the row is zero)
      for (int m = lane; m < n; m += 32) r[m] = len != 0.0 ? (r[m] - mean) / len : 0.0;
    }
    __syncwarp();
    for (int e = lane; e < kRowExt; e += 32) r[n + e] = r[e % n];
  }
  __syncthreads();

  // ---- optimize_cc :488-505: first maximum of the circular cross-correlation of every pair -----------------------
  const int n_groups = (P + 31) / 32, n_tiles = (n + kLagTile - 1) / kLagTile, n_blk = n / kLagTile;
  // units (32 pairs x a share of the lag tiles) are handed out on demand: the warps of a CTA do not progress equally
  // (each unit writes its own result slots, so the outcome does not depend on who takes which)
  for (;;) {
    int u = 0;
    if (lane == 0) u = atomicAdd(sm.flag + 1, 1);
    u = __shfl_sync(0xffffffffu, u, 0);
    if (u >= n_groups * n_split) break;
    const int g = u / n_split, s = u - g * n_split;
    const int tile0 = static_cast<int>(static_cast<long>(n_tiles) * s / n_split);
    const int tile1 = static_cast<int>(static_cast<long>(n_tiles) * (s + 1) / n_split);
    const int p = min(32 * g + lane, P - 1);
    const double* ri = sm.rows + static_cast<size_t>(sm.pi[p]) * stride;
    const double* rj = sm.rows + static_cast<size_t>(sm.pj[p]) * stride;
    double best = -DBL_MAX;
    int best_k = tile0 * kLagTile;
    for (int tile = tile0; tile < tile1; ++tile) {
      const int k0 = tile * kLagTile;
      double acc[kLagTile], W[kLagTile];
#pragma unroll
      for (int t = 0; t < kLagTile / 2; ++t) {
        const double2 v = *reinterpret_cast<const double2*>(rj + k0 + 2 * t);
        W[2 * t] = v.x;
        W[2 * t + 1] = v.y;
        acc[2 * t] = acc[2 * t + 1] = 0.0;
      }
      int base = k0;  // (m + k0) mod n, the same for every lane
      for (int mb = 0; mb < n_blk; ++mb) {
        if (base & 1)
          corr_block<false, true>(ri + mb * kLagTile, rj + base + kLagTile, W, acc, kLagTile);
        else
          corr_block<false, false>(ri + mb * kLagTile, rj + base + kLagTile, W, acc, kLagTile);
        base += kLagTile;
        if (base >= n) base -= n;
      }
      if (n_blk * kLagTile < n) {
        if (base & 1)
          corr_block<true, true>(ri + n_blk * kLagTile, rj + base + kLagTile, W, acc, n - n_blk * kLagTile);
        else
          corr_block<true, false>(ri + n_blk * kLagTile, rj + base + kLagTile, W, acc, n - n_blk * kLagTile);
      }
      if (a.cc && 32 * g + lane < P) {  // correlate: the pair's correlation function of this window, lags k0 .. k0+15
        double* dst = a.cc + (static_cast<size_t>(p) * a.n_win + w) * n + k0;
#pragma unroll
        for (int t = 0; t < kLagTile; ++t)
          if (k0 + t < n) dst[t] = acc[t];
      }
#pragma unroll
      for (int t = 0; t < kLagTile; ++t) {
        if (k0 + t < n && acc[t] > best) {  // strict: maxloc keeps the first maximum
          best = acc[t];
          best_k = k0 + t;
        }
      }
    }
    if (32 * g + lane < P) {
      sm.bestv[static_cast<size_t>(s) * P + p] = best;
      sm.bestk[static_cast<size_t>(s) * P + p] = best_k;
    }
  }
  __syncthreads();
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    double best = sm.bestv[p];
    int best_k = sm.bestk[p];
    for (int s = 1; s < n_split; ++s) {
      const double v = sm.bestv[static_cast<size_t>(s) * P + p];
      if (v > best) {
        best = v;
        best_k = sm.bestk[static_cast<size_t>(s) * P + p];
      }
    }
    sm.lagk[p] = best_k;
    if (a.lag) a.lag[static_cast<size_t>(w) * P + p] = best_k;
    if (a.cc_max) a.cc_max[static_cast<size_t>(p) * a.n_win + w] = best;  // maxval(cc(:, i)), src/cls_correlator.f90:233
  }
  if (a.mode != 0) return;
  __syncthreads();

  // ---- optimize_cc :507-520: station times and their scatter (sums in the reference's order) ---------------------
  double* rel = sm.bestv;  // [P], free from here on
  for (int p = threadIdx.x; p < P; p += blockDim.x) rel[p] = lag_of(sm.lagk[p], n, a.dt);
  __syncthreads();
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    double s = 0.0;
    for (int j = 0; j < S; ++j) s = __dsub_rn(s, antisym(rel, S, i, j));
    sm.tS[i] = s / S;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    const double ti = sm.tS[i];
    double s = 0.0;
    for (int j = 0; j < S; ++j) {
      if (j == i) continue;
      const double d = __dsub_rn(__dsub_rn(sm.tS[j], ti), antisym(rel, S, i, j));
      s = __dadd_rn(s, __dmul_rn(d, d));
    }
    const size_t o = static_cast<size_t>(w) * S + i;
    a.t[o] = ti;
    a.t_stdv[o] = sqrt(s / (S - 2));
    sm.it[i] = static_cast<int>(fmax(-static_cast<double>(n), fmin(static_cast<double>(n), round(ti / a.dt))));  // nint
  }
  __syncthreads();

  // ---- optimize_amp :417-441: the raw window again, shifted by it_i samples ---------------------------------------
  for (int i = warp; i < S; i += n_warps) {
    const double* x = a.env + static_cast<size_t>(i) * a.n_total + j1;
    double* r = sm.rows + static_cast<size_t>(i) * stride;
    for (int m = lane; m < n; m += 32) r[m] = x[m];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    const double* r = sm.rows + static_cast<size_t>(i) * stride;
    const int it = sm.it[i];
    double s = 0.0;
    for (int m = max(0, -it); m < min(n, n - it); ++m) s = __dadd_rn(s, __dmul_rn(r[m + it], r[m + it]));
    sm.sxx[i] = s;
  }
  __syncthreads();
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const int i = sm.pi[p], j = sm.pj[p], ii = sm.it[i], ij = sm.it[j];
    const double* ri = sm.rows + static_cast<size_t>(i) * stride + ii;
    const double* rj = sm.rows + static_cast<size_t>(j) * stride + ij;
    double sxy = 0.0;
    for (int m = max(0, -min(ii, ij)); m < min(n, n - max(ii, ij)); ++m) sxy = __dadd_rn(sxy, __dmul_rn(ri[m], rj[m]));
    if (sxy < 0.0) atomicOr(sm.flag, 1);
    rel[p] = log(sxy / sm.sxx[i]);
  }
  __syncthreads();
  const bool zeroed = *sm.flag != 0;  // :430-434: a negative cross product gives the window up
  // ---- optimize_amp :443-459 -------------------------------------------------------------------------------------
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    double s = 0.0;
    for (int j = 0; j < S; ++j) s = __dsub_rn(s, antisym(rel, S, i, j));
    sm.tS[i] = s / S;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    const double ai = sm.tS[i];
    double s = 0.0;
    for (int j = 0; j < S; ++j) {
      if (j == i) continue;
      const double d = __dsub_rn(__dsub_rn(sm.tS[j], ai), antisym(rel, S, i, j));
      s = __dadd_rn(s, __dmul_rn(d, d));
    }
    const size_t o = static_cast<size_t>(w) * S + i;
    a.amp[o] = zeroed ? 0.0 : ai;
    a.amp_stdv[o] = zeroed ? 0.0 : sqrt(s / (S - 2));
  }
}

// scan_cc (src/cls_measurer.f90:228-258): a window is detected when more than n_pair_thred station pairs have their
// maximum correlation at or above the pair's threshold
__global__ void detect_kernel(const double* __restrict__ cc_max /* [P][n_win] */, const double* __restrict__ thr /* [P][3] */,
                              const int P, const int n_win, const int n_pair_thred, int32_t* __restrict__ detected,
                              int32_t* __restrict__ count) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_win) return;
  int c = 0;
  for (int p = 0; p < P; ++p) c += cc_max[static_cast<size_t>(p) * n_win + w] >= thr[3 * p] ? 1 : 0;
  detected[w] = c > n_pair_thred ? 1 : 0;
  if (count) count[w] = c;
}

cudaError_t launch_detect(const double* cc_max, const double* thr, int P, int n_win, int n_pair_thred, int32_t* detected,
                          int32_t* count, cudaStream_t stream) {
  detect_kernel<<<(n_win + 127) / 128, 128, 0, stream>>>(cc_max, thr, P, n_win, n_pair_thred, detected, count);
  return cudaGetLastError();
}

// 0 = the window does not fit the CTA's shared memory
int measure_split(int S, int n, size_t smem_max, size_t* smem) {
  const int P = S * (S - 1) / 2, n_groups = (P + 31) / 32, n_warps = kMeasureThreads / 32;
  const int n_tiles = (n + kLagTile - 1) / kLagTile;
  int n_split = (kUnitsPerWarp * n_warps + n_groups - 1) / n_groups;  // enough units for the queue to even out the warps
  if (n_split > n_tiles) n_split = n_tiles;
  if (n_split < 1) n_split = 1;
  while (n_split > 1 && measure_smem(S, n, n_split) > smem_max) --n_split;
  *smem = measure_smem(S, n, n_split);
  return *smem <= smem_max ? n_split : 0;
}

cudaError_t launch_measure(const MeasureArgs& a, cudaStream_t stream) {
  int dev = 0, smem_max = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  e = cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e != cudaSuccess) return e;
  size_t smem = 0;
  const int n_split = measure_split(a.S, a.n, static_cast<size_t>(smem_max), &smem);
  if (n_split == 0 || a.S > 65535) return cudaErrorNotSupported;
  e = cudaFuncSetAttribute(measure_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  measure_kernel<<<a.n_win, kMeasureThreads, smem, stream>>>(a, n_split);
  return cudaGetLastError();
}

}  // namespace htm
