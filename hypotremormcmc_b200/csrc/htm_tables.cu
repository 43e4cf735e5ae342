// Device-side build of the observation tables from the driver's raw float64 arrays.
//
// What it computes is init_forward's precomputation (src/cls_forward.f90:71-92) in the table layouts of
// htm_forward.cuh: per (station, event) the precisions w = sigma^-2 and the log sigma terms, with the
// degenerate-sigma rule of :78-90 (the branch looks at t_stdv ONLY; the "else" sets every precision to 1 and
// log sigma := 1.0, not 0), per event the constants C_e = sum_j (0.5 ln 2 pi + ln sigma) over the data types in
// use and 1 / sum_j w.  All arithmetic is float64, sums run over the stations in index order (one thread per
// event), the results are stored in the handle's precision.  The raw arrays arrive by ONE host-to-device copy
// from pinned staging memory, so the host does no per-element work at all.
#include "htm_common.cuh"
#include "htm_kernels.hpp"

namespace htm {

template <typename real>
__global__ void build_tables_kernel(const TableBuild b) {
  typedef typename M<real>::real4 real4;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e == 0) {  // station table (tiny): {X, Y, Z, 0}
    real4* sta = static_cast<real4*>(b.sta4);
    for (int j = 0; j < b.S; ++j) {
      real4 s;
      s.x = static_cast<real>(b.sta_xyz[j]);
      s.y = static_cast<real>(b.sta_xyz[b.S + j]);
      s.z = static_cast<real>(b.sta_xyz[2 * b.S + j]);
      s.w = 0;
      sta[j] = s;
    }
  }
  if (e >= b.E) return;
  const size_t n = static_cast<size_t>(b.E) * b.S;
  const double* t_obs = b.obs_in;
  const double* t_stdv = b.obs_in + n;
  const double* a_obs = b.obs_in + 2 * n;
  const double* a_stdv = b.obs_in + 3 * n;
  // the table with the fixed station terms folded in exists in the factorised mode only
  real4* obs = b.obs4 ? static_cast<real4*>(b.obs4) + static_cast<size_t>(e) * b.S : nullptr;
  real4* raw = static_cast<real4*>(b.obs4_raw) + static_cast<size_t>(e) * b.S;
  double Ce = 0.0, swt = 0.0, swa = 0.0;
  for (int j = 0; j < b.S; ++j) {
    const size_t k = static_cast<size_t>(e) * b.S + j;
    const double ts = t_stdv[k], as = a_stdv[k];
    double wt, wa, lt, la;
    if (ts > 1.e-16) {  // src/cls_forward.f90:78: decided by t_stdv only
      lt = ::log(ts);
      wt = 1.0 / (ts * ts);
      la = ::log(as);
      wa = 1.0 / (as * as);
    } else {  // :84-90
      lt = 1.0;
      wt = 1.0;
      la = 1.0;
      wa = 1.0;
    }
    if (!b.use_time) wt = 0.0;
    if (!b.use_amp) wa = 0.0;
    if (b.use_time) Ce += kLog2PiHalf + lt;
    if (b.use_amp) Ce += kLog2PiHalf + la;
    swt += wt;
    swa += wa;
    real4 r, o;
    r.x = static_cast<real>(t_obs[k]);
    r.y = static_cast<real>(wt);
    r.z = static_cast<real>(a_obs[k]);
    r.w = static_cast<real>(wa);
    o.x = static_cast<real>(t_obs[k] + b.g_tc_ac[j]);
    o.y = r.y;
    o.z = static_cast<real>(a_obs[k] + b.g_tc_ac[b.S + j]);
    o.w = r.w;
    raw[j] = r;
    if (obs) obs[j] = o;
  }
  real4 c;
  c.x = static_cast<real>(Ce);
  c.y = static_cast<real>(swt > 0 ? 1.0 / swt : 0.0);
  c.z = static_cast<real>(swa > 0 ? 1.0 / swa : 0.0);
  c.w = 0;
  static_cast<real4*>(b.evc4)[e] = c;
  const double xm = b.xy_mu ? b.xy_mu[e] : 0.0, ym = b.xy_mu ? b.xy_mu[b.E + e] : 0.0;
  real* pxy = static_cast<real*>(b.prior_xy);
  pxy[2 * e] = static_cast<real>(xm);
  pxy[2 * e + 1] = static_cast<real>(ym);
  if (b.prior_xy64) {
    b.prior_xy64[2 * e] = xm;
    b.prior_xy64[2 * e + 1] = ym;
  }
}

cudaError_t launch_build_tables(int precision, const TableBuild& b, cudaStream_t stream) {
  const unsigned block = 128, grid = static_cast<unsigned>((b.E + block - 1) / block);
  if (precision == HTM_PRECISION_F64)
    build_tables_kernel<double><<<grid, block, 0, stream>>>(b);
  else
    build_tables_kernel<float><<<grid, block, 0, stream>>>(b);
  return cudaGetLastError();
}

// ---- sample rings -> the reference's record layout, on the device ------------------------------------------
// One thread per (record, event): float4/double4 {x, y, z, L_e} of the ring -> three consecutive doubles of
// hypo[record][3*E] (src/hypo_tremor_mcmc.f90:272-274 writes x(1:3E) per record).  Used by the all-gather of
// event-sharded samples, which runs device to device.
template <typename real>
__global__ void pack_hypo_kernel(const typename M<real>::real4* __restrict__ ring, const int* __restrict__ slot_of_rec,
                                 int n_rec, int E, size_t ring_stride /* real4 between consecutive ring records */,
                                 size_t row_offset /* real4 offset of (rank, cold slot) inside a ring record */,
                                 double* __restrict__ out, size_t out_stride /* doubles between output records */) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(n_rec) * E) return;
  const int k = static_cast<int>(i / E), e = static_cast<int>(i % E);
  const typename M<real>::real4 v = ring[static_cast<size_t>(slot_of_rec[k]) * ring_stride + row_offset + e];
  double* o = out + static_cast<size_t>(k) * out_stride + static_cast<size_t>(3) * e;
  o[0] = static_cast<double>(v.x);
  o[1] = static_cast<double>(v.y);
  o[2] = static_cast<double>(v.z);
}

cudaError_t launch_pack_hypo(int precision, const void* ring, const int* slot_of_rec, int n_rec, int E, size_t ring_stride,
                             size_t row_offset, double* out, size_t out_stride, cudaStream_t stream) {
  const size_t n = static_cast<size_t>(n_rec) * E;
  if (n == 0) return cudaSuccess;
  const unsigned block = 256, grid = static_cast<unsigned>((n + block - 1) / block);
  if (precision == HTM_PRECISION_F64)
    pack_hypo_kernel<double><<<grid, block, 0, stream>>>(static_cast<const double4*>(ring), slot_of_rec, n_rec, E, ring_stride,
                                                         row_offset, out, out_stride);
  else
    pack_hypo_kernel<float><<<grid, block, 0, stream>>>(static_cast<const float4*>(ring), slot_of_rec, n_rec, E, ring_stride,
                                                        row_offset, out, out_stride);
  return cudaGetLastError();
}

}  // namespace htm
