// Batched full log-likelihood: forward%calc_log_likelihood (src/cls_forward.f90:268-303) for
// n_models models at once.  Two kernels produce the per-event values, then one block per model adds
// them in a fixed order:
//   loglik_tile_kernel   CTA = tile of 32 events, lane = event; the tile's observation rows arrive by
//                        bulk TMA into padded shared-memory rows and serve up to 8 models; with fewer
//                        than 8 models per CTA the warps also split the stations.  HBM-bound: the table
//                        is read once (16*S B per event in float32) whatever the number of models.
//   loglik_event_kernel  one warp per (model, event), stations across lanes, shuffle reduction: the
//                        fallback when a tile's rows do not fit in shared memory (very large n_sta).
#include <cstdlib>
#include <string>

#include "htm_forward.cuh"
#include "htm_kernels.hpp"

namespace htm {

template <typename real>
__global__ void __launch_bounds__(256) loglik_event_kernel(const typename M<real>::real4* __restrict__ sta4,
                                                           const typename M<real>::real4* __restrict__ obs4,
                                                           const typename M<real>::real4* __restrict__ evc4, int E,
                                                           int S, int n_models, const double* __restrict__ hypo,
                                                           const double* __restrict__ tc, const double* __restrict__ ac,
                                                           const double* __restrict__ vs, const double* __restrict__ qs,
                                                           double* __restrict__ per_event) {
  const long gw = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (gw >= static_cast<long>(n_models) * E) return;
  const int m = static_cast<int>(gw / E), e = static_cast<int>(gw % E);
  const double* h = hypo + (static_cast<size_t>(m) * E + e) * 3;
  const Glob<real> g = make_glob<real>(static_cast<real>(vs[m]), static_cast<real>(qs[m]));
  const real L = warp_event_loglik<real, double>(sta4, obs4 + static_cast<size_t>(e) * S, evc4[e], S,
                                                 static_cast<real>(h[0]), static_cast<real>(h[1]),
                                                 static_cast<real>(h[2]), g, tc + static_cast<size_t>(m) * S,
                                                 ac + static_cast<size_t>(m) * S);
  if ((threadIdx.x & 31) == 0) per_event[gw] = static_cast<double>(L);
}

constexpr int kLTile = 32;   // events per CTA
constexpr int kLWarps = 8;   // warps per CTA = models per CTA x station slices

template <typename real>
static size_t loglik_tile_smem(int S, int mpc) {
  typedef typename M<real>::real4 real4;
  return 16 + static_cast<size_t>(2) * kLTile * (S + 1) * sizeof(real4) + S * sizeof(real4) +
         static_cast<size_t>(2) * mpc * S * sizeof(real) + static_cast<size_t>(2) * kLWarps * 3 * 32 * sizeof(real);
}

// One wave of CTAs; CTA (x, y) walks the tiles x, x + gridDim.x, ... for the models of group y.  The rows of
// tile i+1 are in flight (2-stage TMA ring) while tile i is evaluated, so the kernel streams the table at HBM
// rate instead of paying the load latency once per wave of CTAs.
template <typename real>
__global__ void __launch_bounds__(kLWarps * 32) loglik_tile_kernel(const typename M<real>::real4* __restrict__ sta4,
                                                                   const typename M<real>::real4* __restrict__ obs4,
                                                                   const typename M<real>::real4* __restrict__ evc4, int E,
                                                                   int S, int n_models, int mpc,
                                                                   const double* __restrict__ hypo,
                                                                   const double* __restrict__ tc, const double* __restrict__ ac,
                                                                   const double* __restrict__ vs, const double* __restrict__ qs,
                                                                   double* __restrict__ per_event) {
  typedef typename M<real>::real4 real4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (E + kLTile - 1) / kLTile, row_len = S + 1;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);  // [2]: rows of stage 0 / 1 (the first also carries the station table)
  real4* rows = reinterpret_cast<real4*>(smem_raw + 16);  // [2][kLTile][row_len]
  real4* sta = rows + 2 * kLTile * row_len;
  real* tcs = reinterpret_cast<real*>(sta + S);
  real* acs = tcs + mpc * S;
  real* part = acs + mpc * S;  // [2][kLWarps][3][32]
  const uint32_t row_bytes = static_cast<uint32_t>(S * sizeof(real4));
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
    fence_mbar_init();
    fence_proxy_async();
  }
  __syncthreads();
  auto issue = [&](int tile, int buf, bool with_sta) {  // warp 0
    const int n_ev = min(kLTile, E - tile * kLTile);
    if (lane == 0) mbar_expect_tx(bar + buf, row_bytes * (n_ev + (with_sta ? 1 : 0)));
    __syncwarp();
    if (lane < n_ev)
      tma_load_1d(rows + (buf * kLTile + lane) * row_len, obs4 + static_cast<size_t>(tile * kLTile + lane) * S, row_bytes,
                  bar + buf);
    if (with_sta && lane == 0) tma_load_1d(sta, sta4, row_bytes, bar + buf);
  };
  if (warp == 0) {
    if (static_cast<int>(blockIdx.x) < n_tiles) issue(blockIdx.x, 0, true);
    if (static_cast<int>(blockIdx.x + gridDim.x) < n_tiles) issue(blockIdx.x + gridDim.x, 1, false);
  }
  const int m0 = blockIdx.y * mpc;
  for (int i = threadIdx.x; i < mpc * S; i += blockDim.x) {
    const int mm = i / S, j = i - mm * S, m = min(m0 + mm, n_models - 1);
    tcs[i] = static_cast<real>(tc[static_cast<size_t>(m) * S + j]);
    acs[i] = static_cast<real>(ac[static_cast<size_t>(m) * S + j]);
  }
  __syncthreads();

  const int slices = kLWarps / mpc;  // station slices per model
  const int ml = warp / slices, sl = warp - ml * slices, m = m0 + ml;
  const bool m_ok = m < n_models;
  const real* mt = tcs + ml * S;
  const real* ma = acs + ml * S;
  Glob<real> g = make_glob<real>(static_cast<real>(1), static_cast<real>(1));
  if (m_ok) g = make_glob<real>(static_cast<real>(vs[m]), static_cast<real>(qs[m]));
  // this lane's hypocentre and event constants are fetched one tile ahead (global-memory latency off the
  // critical path of a tile)
  const int mm = m_ok ? m : 0;
  auto event_of = [&](int tile) { return min(tile * kLTile + lane, E - 1); };
  double hx = 0, hy = 0, hz = 0;
  real4 evc = evc4[0];
  if (static_cast<int>(blockIdx.x) < n_tiles) {
    const int e0 = event_of(blockIdx.x);
    const double* h = hypo + (static_cast<size_t>(mm) * E + e0) * 3;
    hx = h[0];
    hy = h[1];
    hz = h[2];
    evc = evc4[e0];
  }
  int it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    const real px = static_cast<real>(hx), py = static_cast<real>(hy), pz = static_cast<real>(hz);
    const real4 evc_now = evc;
    if (tile + static_cast<int>(gridDim.x) < n_tiles) {
      const int e1 = event_of(tile + gridDim.x);
      const double* h = hypo + (static_cast<size_t>(mm) * E + e1) * 3;
      hx = h[0];
      hy = h[1];
      hz = h[2];
      evc = evc4[e1];
    }
    mbar_wait(bar + buf, static_cast<uint32_t>((it >> 1) & 1));
    const int n_ev = min(kLTile, E - tile * kLTile);
    const int e = tile * kLTile + lane;
    const bool ev_ok = e < E;
    const real4* row = rows + (buf * kLTile + (ev_ok ? lane : n_ev - 1)) * row_len;
    real S1t = 0, S1a = 0, S2 = 0;
    if (m_ok) {
      real ct, ca;  // shift: the residuals of station 0 (every slice computes them)
      station_resid(px, py, pz, g, sta[0], row[0], mt[0], ma[0], ct, ca);
#pragma unroll 2
      for (int j = sl; j < S; j += slices) {
        const real4 ob = row[j];
        real rt, ra;
        station_resid(px, py, pz, g, sta[j], ob, mt[j], ma[j], rt, ra);
        const real et = rt - ct, ea = ra - ca;
        const real qt = ob.y * et, qa = ob.w * ea;
        S1t += qt;
        S1a += qa;
        S2 += qt * et;
        S2 += qa * ea;
      }
    }
    real* pt = part + buf * kLWarps * 3 * 32;  // two copies: the combining warps may lag one tile behind
    if (slices > 1) {
      pt[(warp * 3 + 0) * 32 + lane] = S1t;
      pt[(warp * 3 + 1) * 32 + lane] = S1a;
      pt[(warp * 3 + 2) * 32 + lane] = S2;
    }
    __syncthreads();  // every warp is done with stage `buf`: refill it with the tile after next
    if (warp == 0 && tile + 2 * static_cast<int>(gridDim.x) < n_tiles) issue(tile + 2 * gridDim.x, buf, false);
    if (sl == 0) {
      for (int q = 1; q < slices; ++q) {  // fixed order
        S1t += pt[((warp + q) * 3 + 0) * 32 + lane];
        S1a += pt[((warp + q) * 3 + 1) * 32 + lane];
        S2 += pt[((warp + q) * 3 + 2) * 32 + lane];
      }
      if (m_ok && ev_ok)
        per_event[static_cast<size_t>(m) * E + e] =
            static_cast<double>(finish_loglik<real>(S1t, S2, S1a, static_cast<real>(0), evc_now));
    }
  }
}

// fixed-order sum of per_event[m][0..E): strided partials, then a shared-memory tree
__global__ void __launch_bounds__(256) loglik_reduce_kernel(const double* __restrict__ per_event, int E,
                                                            double* __restrict__ L) {
  __shared__ double s[256];
  const int m = blockIdx.x;
  double acc = 0.0;
  for (int e = threadIdx.x; e < E; e += 256) acc += per_event[static_cast<size_t>(m) * E + e];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) L[m] = s[0];
}

cudaError_t launch_loglik(int precision, const Tables& tab, int E, int S, int M_, const double* hypo,
                          const double* tc, const double* ac, const double* vs, const double* qs,
                          double* per_event, double* L, cudaStream_t stream) {
  // HTM_LOGLIK_KERNEL=warp|tile pins the kernel (tests)
  const char* pin = std::getenv("HTM_LOGLIK_KERNEL");
  const int mpc = M_ >= 8 ? 8 : (M_ >= 4 ? 4 : (M_ >= 2 ? 2 : 1));
  const size_t smem = precision == HTM_PRECISION_F64 ? loglik_tile_smem<double>(S, mpc) : loglik_tile_smem<float>(S, mpc);
  if (smem <= 200 * 1024 && !(pin && std::string(pin) == "warp")) {
    // one wave of CTAs per model group, each walking its tiles through the TMA ring
    const int n_tiles = (E + kLTile - 1) / kLTile, gy = (M_ + mpc - 1) / mpc;
    int dev = 0, n_sm = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e2;
    if (precision == HTM_PRECISION_F64) {
      e2 = cudaFuncSetAttribute(loglik_tile_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e2 == cudaSuccess)
        e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, loglik_tile_kernel<double>, kLWarps * 32, smem);
    } else {
      e2 = cudaFuncSetAttribute(loglik_tile_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e2 == cudaSuccess)
        e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, loglik_tile_kernel<float>, kLWarps * 32, smem);
    }
    if (e2 != cudaSuccess) return e2;
    int gx = per_sm * n_sm / gy;
    if (gx < 1) gx = 1;
    if (gx > n_tiles) gx = n_tiles;
    const dim3 tgrid(gx, gy);
    if (precision == HTM_PRECISION_F64) {
      loglik_tile_kernel<double><<<tgrid, kLWarps * 32, smem, stream>>>(
          static_cast<const double4*>(tab.sta4), static_cast<const double4*>(tab.obs4_raw),
          static_cast<const double4*>(tab.evc4), E, S, M_, mpc, hypo, tc, ac, vs, qs, per_event);
    } else {
      loglik_tile_kernel<float><<<tgrid, kLWarps * 32, smem, stream>>>(
          static_cast<const float4*>(tab.sta4), static_cast<const float4*>(tab.obs4_raw),
          static_cast<const float4*>(tab.evc4), E, S, M_, mpc, hypo, tc, ac, vs, qs, per_event);
    }
    e2 = cudaGetLastError();
    if (e2 != cudaSuccess) return e2;
    loglik_reduce_kernel<<<M_, 256, 0, stream>>>(per_event, E, L);
    return cudaGetLastError();
  }
  const long n_warps = static_cast<long>(M_) * E;
  const unsigned block = 256;
  const unsigned grid = static_cast<unsigned>((n_warps * 32 + block - 1) / block);
  if (precision == HTM_PRECISION_F64) {
    loglik_event_kernel<double><<<grid, block, 0, stream>>>(
        static_cast<const double4*>(tab.sta4), static_cast<const double4*>(tab.obs4_raw),
        static_cast<const double4*>(tab.evc4), E, S, M_, hypo, tc, ac, vs, qs, per_event);
  } else {
    loglik_event_kernel<float><<<grid, block, 0, stream>>>(
        static_cast<const float4*>(tab.sta4), static_cast<const float4*>(tab.obs4_raw),
        static_cast<const float4*>(tab.evc4), E, S, M_, hypo, tc, ac, vs, qs, per_event);
  }
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return err;
  loglik_reduce_kernel<<<M_, 256, 0, stream>>>(per_event, E, L);
  return cudaGetLastError();
}

}  // namespace htm
