// TEST INFRASTRUCTURE -- CPU restatement of the upstream QC stage hypo_tremor_select, line by line:
//   src/cls_selector.f90:61-67   distance table d(receiver, source) with the assumed source depth z_guess
//   src/cls_selector.f90:75-132  eval_wave_propagation: nearest station = maxloc(a) (first maximum), geometrical
//                                spreading correction a += log d, two weighted fits, two correlation coefficients
//   src/mod_regress.f90:5-38     linear_regression: sums accumulated left to right, a = slope, b = intercept
//   src/mod_regress.f90:40-58    weighted_corr: weighted means, UNWEIGHTED sums of squares and products
//   src/hypo_tremor_select.f90:122-127  acceptance window on vs and B
// Parity unpinned by the reference (no tests, no Fortran compiler here): pinned by an independent numpy restatement
// and an analytic known answer in tests/test_select.py.  Never linked by the product.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace hto {

inline void linear_regression(int n, const double* x, const double* y, const double* w, double* a, double* b) {
  double sumx = 0.0, sumy = 0.0, sumw = 0.0, sumxy = 0.0, sumx2 = 0.0;
  for (int i = 0; i < n; ++i) {
    sumx = sumx + x[i] * w[i];
    sumy = sumy + y[i] * w[i];
    sumw = sumw + w[i];
    sumxy = sumxy + x[i] * y[i] * w[i];
    sumx2 = sumx2 + x[i] * x[i] * w[i];
  }
  const double d = sumw * sumx2 - sumx * sumx;
  *a = (sumw * sumxy - sumx * sumy) / d;
  *b = (sumx2 * sumy - sumx * sumxy) / d;
}

inline double weighted_corr(int n, const double* x, const double* y, const double* w) {
  double sum_w = 0.0, sx = 0.0, sy = 0.0;
  for (int i = 0; i < n; ++i) sum_w = sum_w + w[i];
  for (int i = 0; i < n; ++i) sx = sx + x[i] * w[i];
  for (int i = 0; i < n; ++i) sy = sy + y[i] * w[i];
  const double mean_x = sx / sum_w, mean_y = sy / sum_w;
  double s_xx = 0.0, s_yy = 0.0, s_xy = 0.0;
  for (int i = 0; i < n; ++i) s_xx = s_xx + (x[i] - mean_x) * (x[i] - mean_x);
  for (int i = 0; i < n; ++i) s_yy = s_yy + (y[i] - mean_y) * (y[i] - mean_y);
  for (int i = 0; i < n; ++i) s_xy = s_xy + (x[i] - mean_x) * (y[i] - mean_y);
  return s_xy / std::sqrt(s_xx * s_yy);
}

// one window: t, t_err, a, a_err [S] -> vs, t0, b, a0, cc_t, cc_a
inline void select_window(int S, const double* sta_x, const double* sta_y, const double* sta_z, double z_guess,
                          const double* t, const double* t_err, const double* a_in, const double* a_err, double out[6]) {
  int near = 0;  // maxloc: the first maximum
  for (int j = 1; j < S; ++j)
    if (a_in[j] > a_in[near]) near = j;
  std::vector<double> d(S), a(S), w_t(S), w_a(S);
  for (int j = 0; j < S; ++j) {
    const double dx = sta_x[j] - sta_x[near], dy = sta_y[j] - sta_y[near], dz = sta_z[j] - z_guess;
    d[j] = std::sqrt(dx * dx + dy * dy + dz * dz);
    a[j] = a_in[j] + std::log(d[j]);
    w_t[j] = 1.0 / (t_err[j] * t_err[j]);
    w_a[j] = 1.0 / (a_err[j] * a_err[j]);
  }
  double slope, intercept;
  linear_regression(S, d.data(), t, w_t.data(), &slope, &intercept);
  out[0] = 1.0 / slope;
  out[1] = intercept;
  linear_regression(S, d.data(), a.data(), w_a.data(), &slope, &intercept);
  out[2] = -1.0 * slope;
  out[3] = intercept;
  out[4] = weighted_corr(S, d.data(), t, w_t.data());
  out[5] = weighted_corr(S, d.data(), a.data(), w_a.data());
}

}  // namespace hto
