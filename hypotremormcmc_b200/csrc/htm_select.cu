// Upstream QC stage `hypo_tremor_select`, batched over all detected windows.
//
// What the reference does per window (src/cls_selector.f90:75-132, src/mod_regress.f90:5-58; the driver
// src/hypo_tremor_select.f90:84-133 loops over windows, one MPI rank per block of them): take the station with the
// largest log-amplitude as the source's surface projection (maxloc: the first maximum), use its distance table
// d(j) = |X_j - (X_near, Y_near, z_guess)| (src/cls_selector.f90:61-67: the STATION depth enters, not the
// difference of depths -- kept), correct the amplitudes for geometrical spreading (a += ln d), then two weighted
// straight-line fits against distance -- arrival time (slope = 1/vs, intercept t0) and amplitude (slope = -B,
// intercept a0), weights 1/err^2 -- and the correlation coefficients of mod_regress' weighted_corr (weighted
// means, UNWEIGHTED sums of squares -- kept).  A window is selected when vs_min <= vs <= vs_max and
// b_min <= B <= b_max (src/hypo_tremor_select.f90:122-127).
//
// Mapping: one warp per window.  The four observation rows of the window (t, t_err, a, a_err; 32 S bytes) are
// read coalesced into shared memory, lane 0 then runs the reference's loops in station order in float64, so the
// sums round as the reference's do.  The kernel is HBM-bound: 32 S bytes in, 52 bytes out per window.
#include "htm_common.cuh"
#include "htm_kernels.hpp"

namespace htm {

__global__ void __launch_bounds__(256) select_kernel(const SelectArgs a) {
  extern __shared__ double s_sel[];  // per warp: t, t_err, a, a_err, d  [5][S]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int e = blockIdx.x * wpb + warp;
  if (e >= a.E) return;
  const int S = a.S;
  double* t = s_sel + static_cast<size_t>(warp) * 5 * S;
  double* te = t + S;
  double* am = te + S;
  double* ae = am + S;
  double* d = ae + S;
  const size_t o = static_cast<size_t>(e) * S;
  for (int j = lane; j < S; j += 32) {
    t[j] = a.t[o + j];
    te[j] = a.t_err[o + j];
    am[j] = a.a[o + j];
    ae[j] = a.a_err[o + j];
  }
  __syncwarp();
  if (lane != 0) return;
  // maxloc(a): the first maximum
  int near = 0;
  for (int j = 1; j < S; ++j)
    if (am[j] > am[near]) near = j;
  const double xs = a.sta_x[near], ys = a.sta_y[near];
  for (int j = 0; j < S; ++j) {
    const double dx = a.sta_x[j] - xs, dy = a.sta_y[j] - ys, dz = a.sta_z[j] - a.z_guess;
    d[j] = ::sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
    am[j] = am[j] + ::log(d[j]);  // geometrical spreading correction
  }
  // linear_regression (src/mod_regress.f90:5-38), no FMA contraction: the reference's default build has none
  auto regress = [&](const double* y, const double* err, double& slope, double& icpt) {
    double sumx = 0, sumy = 0, sumw = 0, sumxy = 0, sumx2 = 0;
    for (int j = 0; j < S; ++j) {
      const double w = 1.0 / __dmul_rn(err[j], err[j]);
      sumx = __dadd_rn(sumx, __dmul_rn(d[j], w));
      sumy = __dadd_rn(sumy, __dmul_rn(y[j], w));
      sumw = __dadd_rn(sumw, w);
      sumxy = __dadd_rn(sumxy, __dmul_rn(__dmul_rn(d[j], y[j]), w));
      sumx2 = __dadd_rn(sumx2, __dmul_rn(__dmul_rn(d[j], d[j]), w));
    }
    const double den = __dadd_rn(__dmul_rn(sumw, sumx2), -__dmul_rn(sumx, sumx));
    slope = __dadd_rn(__dmul_rn(sumw, sumxy), -__dmul_rn(sumx, sumy)) / den;
    icpt = __dadd_rn(__dmul_rn(sumx2, sumy), -__dmul_rn(sumx, sumxy)) / den;
  };
  // weighted_corr (src/mod_regress.f90:40-58)
  auto corr = [&](const double* y, const double* err) -> double {
    double sw = 0, sx = 0, sy = 0;
    for (int j = 0; j < S; ++j) sw = __dadd_rn(sw, 1.0 / __dmul_rn(err[j], err[j]));
    for (int j = 0; j < S; ++j) sx = __dadd_rn(sx, __dmul_rn(d[j], 1.0 / __dmul_rn(err[j], err[j])));
    for (int j = 0; j < S; ++j) sy = __dadd_rn(sy, __dmul_rn(y[j], 1.0 / __dmul_rn(err[j], err[j])));
    const double mx = sx / sw, my = sy / sw;
    double sxx = 0, syy = 0, sxy = 0;
    for (int j = 0; j < S; ++j) sxx = __dadd_rn(sxx, __dmul_rn(d[j] - mx, d[j] - mx));
    for (int j = 0; j < S; ++j) syy = __dadd_rn(syy, __dmul_rn(y[j] - my, y[j] - my));
    for (int j = 0; j < S; ++j) sxy = __dadd_rn(sxy, __dmul_rn(d[j] - mx, y[j] - my));
    return sxy / ::sqrt(__dmul_rn(sxx, syy));
  };
  double slope, icpt;
  regress(t, te, slope, icpt);
  const double vs = 1.0 / slope, t0 = icpt;
  regress(am, ae, slope, icpt);
  const double b = -1.0 * slope, a0 = icpt;
  a.vs[e] = vs;
  a.t0[e] = t0;
  a.b[e] = b;
  a.a0[e] = a0;
  a.cc_t[e] = corr(t, te);
  a.cc_a[e] = corr(am, ae);
  a.selected[e] = (vs >= a.vs_min && vs <= a.vs_max && b >= a.b_min && b <= a.b_max) ? 1 : 0;
}

cudaError_t launch_select(const SelectArgs& a, cudaStream_t stream) {
  int wpb = 8;
  while (wpb > 1 && static_cast<size_t>(wpb) * 5 * a.S * sizeof(double) > 96 * 1024) wpb >>= 1;
  const size_t smem = static_cast<size_t>(wpb) * 5 * a.S * sizeof(double);
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  cudaError_t err = cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (err != cudaSuccess) return err;
  select_kernel<<<static_cast<unsigned>((a.E + wpb - 1) / wpb), wpb * 32, smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace htm
