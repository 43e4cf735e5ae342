// Batched full log-likelihood: forward%calc_log_likelihood (src/cls_forward.f90:268-303) for
// n_models models at once.  One warp per (model, event): stations across lanes, shuffle
// reduction; then one block per model adds the per-event values in a fixed order.
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
