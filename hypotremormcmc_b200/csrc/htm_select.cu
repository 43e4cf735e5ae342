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
// Mapping: one THREAD per window (a warp-per-window version with lane 0 running the loops issued 32 times the
// instructions: 1.95 ms for 100 000 x 50, 1.3 % of the HBM roofline).  Each thread runs the reference's loops in
// station order in float64, so the sums round as the reference's do; its distance row d(j) and corrected amplitudes
// live in shared memory (thread-minor layout: conflict-free), the observation rows are re-read through L1 (a
// thread consumes whole 32-byte sectors over four consecutive stations).  HBM traffic: 32 S bytes in, 52 out per
// window.
#include "htm_common.cuh"
#include "htm_kernels.hpp"

namespace htm {

constexpr int kSelThreads = 128;

__global__ void __launch_bounds__(kSelThreads) select_kernel(const SelectArgs a) {
  extern __shared__ double s_sel[];  // [2][S][blockDim]: d, corrected amplitude
  const int e = blockIdx.x * blockDim.x + threadIdx.x, S = a.S, nt = blockDim.x;
  if (e >= a.E) return;
  double* d = s_sel + threadIdx.x;
  double* ac = s_sel + static_cast<size_t>(S) * nt + threadIdx.x;
  const size_t o = static_cast<size_t>(e) * S;
  const double* t = a.t + o;
  const double* te = a.t_err + o;
  const double* am = a.a + o;
  const double* ae = a.a_err + o;
  // maxloc(a): the first maximum
  int near = 0;
  double amax = am[0];
  for (int j = 1; j < S; ++j) {
    const double v = am[j];
    if (v > amax) {
      amax = v;
      near = j;
    }
  }
  const double xs = a.sta_x[near], ys = a.sta_y[near];
  for (int j = 0; j < S; ++j) {
    const double dx = a.sta_x[j] - xs, dy = a.sta_y[j] - ys, dz = a.sta_z[j] - a.z_guess;
    const double dj = ::sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
    d[static_cast<size_t>(j) * nt] = dj;
    ac[static_cast<size_t>(j) * nt] = am[j] + ::log(dj);  // geometrical spreading correction
  }
  // linear_regression (src/mod_regress.f90:5-38), no FMA contraction: the reference's default build has none
  auto regress = [&](const double* y, const size_t ys_, const double* err, double& slope, double& icpt) {
    double sumx = 0, sumy = 0, sumw = 0, sumxy = 0, sumx2 = 0;
    for (int j = 0; j < S; ++j) {
      const double w = 1.0 / __dmul_rn(err[j], err[j]), x = d[static_cast<size_t>(j) * nt], yj = y[j * ys_];
      sumx = __dadd_rn(sumx, __dmul_rn(x, w));
      sumy = __dadd_rn(sumy, __dmul_rn(yj, w));
      sumw = __dadd_rn(sumw, w);
      sumxy = __dadd_rn(sumxy, __dmul_rn(__dmul_rn(x, yj), w));
      sumx2 = __dadd_rn(sumx2, __dmul_rn(__dmul_rn(x, x), w));
    }
    const double den = __dadd_rn(__dmul_rn(sumw, sumx2), -__dmul_rn(sumx, sumx));
    slope = __dadd_rn(__dmul_rn(sumw, sumxy), -__dmul_rn(sumx, sumy)) / den;
    icpt = __dadd_rn(__dmul_rn(sumx2, sumy), -__dmul_rn(sumx, sumxy)) / den;
  };
  // weighted_corr (src/mod_regress.f90:40-58): weighted means, unweighted sums of squares
  auto corr = [&](const double* y, const size_t ys_, const double* err) -> double {
    double sw = 0, sx = 0, sy = 0;
    for (int j = 0; j < S; ++j) {
      const double w = 1.0 / __dmul_rn(err[j], err[j]);
      sw = __dadd_rn(sw, w);
      sx = __dadd_rn(sx, __dmul_rn(d[static_cast<size_t>(j) * nt], w));
      sy = __dadd_rn(sy, __dmul_rn(y[j * ys_], w));
    }
    const double mx = sx / sw, my = sy / sw;
    double sxx = 0, syy = 0, sxy = 0;
    for (int j = 0; j < S; ++j) {
      const double ex = d[static_cast<size_t>(j) * nt] - mx, ey = y[j * ys_] - my;
      sxx = __dadd_rn(sxx, __dmul_rn(ex, ex));
      syy = __dadd_rn(syy, __dmul_rn(ey, ey));
      sxy = __dadd_rn(sxy, __dmul_rn(ex, ey));
    }
    return sxy / ::sqrt(__dmul_rn(sxx, syy));
  };
  double slope, icpt;
  regress(t, 1, te, slope, icpt);
  const double vs = 1.0 / slope, t0 = icpt;
  regress(ac, nt, ae, slope, icpt);
  const double b = -1.0 * slope, a0 = icpt;
  a.vs[e] = vs;
  a.t0[e] = t0;
  a.b[e] = b;
  a.a0[e] = a0;
  a.cc_t[e] = corr(t, 1, te);
  a.cc_a[e] = corr(ac, nt, ae);
  a.selected[e] = (vs >= a.vs_min && vs <= a.vs_max && b >= a.b_min && b <= a.b_max) ? 1 : 0;
}

cudaError_t launch_select(const SelectArgs& a, cudaStream_t stream) {
  int nt = kSelThreads;
  while (nt > 32 && static_cast<size_t>(2) * a.S * nt * sizeof(double) > 100 * 1024) nt >>= 1;
  const size_t smem = static_cast<size_t>(2) * a.S * nt * sizeof(double);
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  cudaError_t err = cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (err != cudaSuccess) return err;
  select_kernel<<<static_cast<unsigned>((a.E + nt - 1) / nt), nt, smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace htm
