// Device-side posterior summary: what `hypo_tremor_statistics` computes from the .out files
// (src/cls_statistics.f90:216-264 hypo.stat, :345-431 station_corrections.stat / uniform_structure.stat) --
// per marginal the sorted sample at the 1-based indices il = 0.025 n, im = 0.5 n, iu = 0.975 n (:230-232,
// default-real arithmetic truncated to integer) -- without the samples ever leaving the GPU.
//
// Every htm_run appends the post-burn-in cold-chain records it produced to a marginal-major store
// (store[marginal][sample]); htm_posterior_quantiles then runs ONE exact selection per marginal: a CTA finds the
// three order statistics by most-significant-digit radix selection on order-preserving integer keys (8 bits per
// pass, histogram in shared memory, the marginal's samples streamed from L2/HBM once per pass) -- no sort, and
// bit-for-bit the values a full sort would put at those positions.
#include "htm_common.cuh"
#include "htm_kernels.hpp"

namespace htm {

// ---- append: ring records -> marginal-major store -----------------------------------------------------------
// ring: real4 {x, y, z, L_e} rows of E events; row r of the ring = (slot, rank, cold chain) in the factorised mode,
// (slot, cold slot) in the blocked-Gibbs mode.  The samples appended by one htm_run are consecutive ring rows
// row0 .. row0 + n_new - 1 (the post-burn-in records are a suffix of the slots the run filled).
template <typename real>
__global__ void store_append_hypo_kernel(const typename M<real>::real4* __restrict__ ring, const int row0,
                                         int n_new, int E, real* __restrict__ store, size_t cap, size_t pos0) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(n_new) * E) return;
  const int k = static_cast<int>(i / E), e = static_cast<int>(i % E);
  const typename M<real>::real4 v = ring[static_cast<size_t>(row0 + k) * E + e];
  real* o = store + static_cast<size_t>(3) * e * cap + pos0 + k;
  o[0] = v.x;
  o[cap] = v.y;
  o[2 * cap] = v.z;
}
// shared parameters of the blocked-Gibbs mode: marginals vs, qs, t_corr[S], a_corr[S] (float64 records)
__global__ void store_append_shared_kernel(const double* __restrict__ rec_vs, const double* __restrict__ rec_qs,
                                           const double* __restrict__ rec_tc, const double* __restrict__ rec_ac,
                                           const int row0, int n_new, int S, double* __restrict__ store,
                                           size_t cap, size_t pos0) {
  const int n_marg = 2 + 2 * S;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(n_new) * n_marg) return;
  const int k = static_cast<int>(i / n_marg), q = static_cast<int>(i % n_marg);
  const size_t r = static_cast<size_t>(row0) + k;
  double v;
  if (q == 0) v = rec_vs[r];
  else if (q == 1) v = rec_qs[r];
  else if (q < 2 + S) v = rec_tc[r * S + (q - 2)];
  else v = rec_ac[r * S + (q - 2 - S)];
  store[static_cast<size_t>(q) * cap + pos0 + k] = v;
}

cudaError_t launch_store_append_hypo(int precision, const void* ring, int row0, int n_new, int E, void* store,
                                     size_t cap, size_t pos0, cudaStream_t stream) {
  const size_t n = static_cast<size_t>(n_new) * E;
  if (n == 0) return cudaSuccess;
  const unsigned block = 256, grid = static_cast<unsigned>((n + block - 1) / block);
  if (precision == HTM_PRECISION_F64)
    store_append_hypo_kernel<double><<<grid, block, 0, stream>>>(static_cast<const double4*>(ring), row0, n_new, E,
                                                                 static_cast<double*>(store), cap, pos0);
  else
    store_append_hypo_kernel<float><<<grid, block, 0, stream>>>(static_cast<const float4*>(ring), row0, n_new, E,
                                                                static_cast<float*>(store), cap, pos0);
  return cudaGetLastError();
}
cudaError_t launch_store_append_shared(const double* rec_vs, const double* rec_qs, const double* rec_tc, const double* rec_ac,
                                       int row0, int n_new, int S, double* store, size_t cap, size_t pos0,
                                       cudaStream_t stream) {
  const size_t n = static_cast<size_t>(n_new) * (2 + 2 * S);
  if (n == 0) return cudaSuccess;
  store_append_shared_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(rec_vs, rec_qs, rec_tc, rec_ac, row0,
                                                                                           n_new, S, store, cap, pos0);
  return cudaGetLastError();
}

// ---- exact order statistics by radix selection ------------------------------------------------------------------
template <typename T>
struct Key;
template <>
struct Key<float> {
  typedef uint32_t type;
  static constexpr int kBits = 32;
  static __device__ __forceinline__ uint32_t of(float v) {
    const uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // order-preserving: negative values reversed below the positive ones
  }
  static __device__ __forceinline__ float back(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
  }
};
template <>
struct Key<double> {
  typedef unsigned long long type;
  static constexpr int kBits = 64;
  static __device__ __forceinline__ unsigned long long of(double v) {
    const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v));
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
  }
  static __device__ __forceinline__ double back(unsigned long long k) {
    return __longlong_as_double(static_cast<long long>((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
  }
};

// one CTA per marginal; ranks[3] are 0-based positions in sorted order; out[marginal][3] as double.
// Most-significant-digit radix selection, 8 bits per pass.  As soon as the bucket that holds the wanted rank has at
// most kCand keys, they are gathered into shared memory (one more scan) and the remaining digits are resolved there:
// 3-4 scans of the marginal instead of 8 for float64 (htm_detect_windows selects among 1e5-1e7 values per pair).
constexpr int kCand = 4096;
template <typename T>
__global__ void __launch_bounds__(256) quantile_select_kernel(const T* __restrict__ store, size_t cap, int n, int r0, int r1,
                                                              int r2, double* __restrict__ out) {
  typedef typename Key<T>::type key_t;
  __shared__ unsigned int hist[256];
  __shared__ unsigned int s_bin, s_below, s_count, s_ncand;
  __shared__ key_t cand[kCand];
  const T* x = store + static_cast<size_t>(blockIdx.x) * cap;
  const int ranks[3] = {r0, r1, r2};
  for (int q = 0; q < 3; ++q) {
    if (q > 0 && ranks[q] == ranks[q - 1]) {  // a repeated rank: the same element
      if (threadIdx.x == 0) out[static_cast<size_t>(blockIdx.x) * 3 + q] = out[static_cast<size_t>(blockIdx.x) * 3 + q - 1];
      continue;
    }
    key_t prefix = 0, mask = 0;
    unsigned int rank = static_cast<unsigned int>(ranks[q]);  // rank among the keys that match the prefix so far
    bool gathered = false;
    int n_cand = 0;
    for (int shift = Key<T>::kBits - 8; shift >= 0; shift -= 8) {
      hist[threadIdx.x] = 0;
      __syncthreads();
      if (gathered) {
        for (int i = threadIdx.x; i < n_cand; i += blockDim.x) {
          const key_t k = cand[i];
          if ((k & mask) == prefix) atomicAdd(&hist[static_cast<unsigned int>((k >> shift) & 0xff)], 1u);
        }
      } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
          const key_t k = Key<T>::of(x[i]);
          if ((k & mask) == prefix) atomicAdd(&hist[static_cast<unsigned int>((k >> shift) & 0xff)], 1u);
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) {  // the digit whose bucket holds position `rank`
        unsigned int below = 0, b = 0;
        for (; b < 255; ++b) {
          if (below + hist[b] > rank) break;
          below += hist[b];
        }
        s_bin = b;
        s_below = below;
        s_count = hist[b];
        s_ncand = 0;
      }
      __syncthreads();
      prefix |= static_cast<key_t>(s_bin) << shift;
      mask |= static_cast<key_t>(0xff) << shift;
      rank -= s_below;
      const unsigned int in_bucket = s_count;
      __syncthreads();
      if (!gathered && shift > 0 && in_bucket <= static_cast<unsigned int>(kCand)) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
          const key_t k = Key<T>::of(x[i]);
          if ((k & mask) == prefix) cand[atomicAdd(&s_ncand, 1u)] = k;
        }
        __syncthreads();
        n_cand = static_cast<int>(s_ncand);
        gathered = true;
      }
    }
    if (threadIdx.x == 0) out[static_cast<size_t>(blockIdx.x) * 3 + q] = static_cast<double>(Key<T>::back(prefix));
    __syncthreads();
  }
}

cudaError_t launch_quantile_select(int precision_bits, const void* store, size_t cap, int n_marginals, int n, int r0, int r1,
                                   int r2, double* out, cudaStream_t stream) {
  if (n_marginals == 0) return cudaSuccess;
  if (precision_bits == 64)
    quantile_select_kernel<double><<<n_marginals, 256, 0, stream>>>(static_cast<const double*>(store), cap, n, r0, r1, r2, out);
  else
    quantile_select_kernel<float><<<n_marginals, 256, 0, stream>>>(static_cast<const float*>(store), cap, n, r0, r1, r2, out);
  return cudaGetLastError();
}

}  // namespace htm
