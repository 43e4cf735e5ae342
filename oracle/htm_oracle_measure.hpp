// TEST INFRASTRUCTURE -- CPU restatement of the numerical core of the upstream stage hypo_tremor_measure:
//   src/mod_signal_process.f90:10-28  apply_taper: 5 % cosine taper on both ends (nleng = int(n * 0.05))
//   src/cls_measurer.f90:466-523      optimize_cc: per station, taper and divide by the window's sum of squares;
//                                     per station pair i < j the circular cross-correlation, its FIRST maximum
//                                     (maxloc) turned into a signed lag; t_i = -(1/S) sum_j lag(i, j) and the
//                                     scatter sqrt(sum_{j/=i} (t_j - t_i - lag(i, j))^2 / (S - 2))
//   src/cls_measurer.f90:405-462      optimize_amp: envelopes shifted by nint(t_i / dt) samples (zero outside the
//                                     window), log(sxy / sxx_i) per pair, the same averaging and scatter; any negative
//                                     cross product zeroes the whole window's amplitudes
//   src/cls_measurer.f90:317-400      measure_lag_time: window id -> samples (id - 1) * n_step + 1 ... + n of each
//                                     station's merged envelope; optimize_cc then optimize_amp on the raw window
// The reference forms the correlation with FFTW (r2c of every station, conjg(X_i) * X_j, unnormalised c2r, :478-494).
// That evaluates exactly r(k) = n * sum_m a_i(m) a_j((m + k) mod n); the restatement below is that sum in float64
// (FFTW is not in this image).  tests/test_measure.py holds the transform route in numpy and pins this file to it.
// Parity unpinned by the reference (no tests, no Fortran compiler, no FFTW here).  Never linked by the product.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace hto {

inline void apply_taper(int n, const double* x, double* out) {
  const double pi = std::acos(-1.0);
  const int nleng = static_cast<int>(n * 0.05);
  for (int i = 0; i < n; ++i) out[i] = x[i];
  for (int i = 1; i <= nleng; ++i) {
    const double fac = 0.5 * (1.0 - std::cos((i - 1) * pi / nleng));
    out[i - 1] = x[i - 1] * fac;
    out[n - i] = x[n - i] * fac;
  }
}

// pair (i, j), i < j, in the reference's loop order -> 0-based position
inline int pair_index(int S, int i, int j) { return i * (2 * S - i - 1) / 2 + (j - i - 1); }

// x [S][n]; t, t_stdv [S]; lag_k (may be null) [S (S - 1) / 2]: 0-based sample index of the correlation maximum
inline void optimize_cc(int S, int n, double dt, const double* x, double* t, double* t_stdv, int32_t* lag_k) {
  std::vector<double> a(static_cast<size_t>(S) * n), lag_t(static_cast<size_t>(S) * S, 0.0), r(n);
  for (int i = 0; i < S; ++i) {
    const double* xi = x + static_cast<size_t>(i) * n;
    double* ai = a.data() + static_cast<size_t>(i) * n;
    double l = 0.0;
    for (int m = 0; m < n; ++m) l = l + xi[m] * xi[m];
    apply_taper(n, xi, ai);
    for (int m = 0; m < n; ++m) ai[m] = ai[m] / l;
  }
  for (int i = 0; i < S - 1; ++i) {
    for (int j = i + 1; j < S; ++j) {
      const double *ai = a.data() + static_cast<size_t>(i) * n, *aj = a.data() + static_cast<size_t>(j) * n;
      int best = 0;
      for (int k = 0; k < n; ++k) {
        double s = 0.0;
        for (int m = 0; m < n - k; ++m) s = s + ai[m] * aj[m + k];
        for (int m = n - k; m < n; ++m) s = s + ai[m] * aj[m + k - n];
        r[k] = s;
        if (r[k] > r[best]) best = k;  // maxloc: the first maximum
      }
      const int ilag = best + 1;
      const double lag = ilag <= n / 2 ? (ilag - 1) * dt : (ilag - n - 1) * dt;
      lag_t[static_cast<size_t>(i) * S + j] = lag;
      lag_t[static_cast<size_t>(j) * S + i] = -lag;
      if (lag_k) lag_k[pair_index(S, i, j)] = best;
    }
  }
  for (int i = 0; i < S; ++i) {
    double s = 0.0;
    for (int j = 0; j < S; ++j) s = s - lag_t[static_cast<size_t>(i) * S + j];
    t[i] = s / S;
  }
  for (int i = 0; i < S; ++i) {
    double s = 0.0;
    for (int j = 0; j < S; ++j) {
      if (i == j) continue;
      const double d = t[j] - t[i] - lag_t[static_cast<size_t>(i) * S + j];
      s = s + d * d;
    }
    t_stdv[i] = std::sqrt(s / (S - 2));
  }
}

inline void optimize_amp(int S, int n, double dt, const double* x, const double* t, double* amp, double* amp_stdv) {
  std::vector<double> x2(static_cast<size_t>(S) * n, 0.0), rel(static_cast<size_t>(S) * S, 0.0), sxx(S);
  for (int i = 0; i < S; ++i) {
    const long it = std::lround(t[i] / dt);  // nint: halves away from zero
    double s = 0.0;
    for (int j = 0; j < n; ++j) {
      const long src = j + it;
      if (src >= 0 && src < n) x2[static_cast<size_t>(i) * n + j] = x[static_cast<size_t>(i) * n + src];
    }
    for (int j = 0; j < n; ++j) s = s + x2[static_cast<size_t>(i) * n + j] * x2[static_cast<size_t>(i) * n + j];
    sxx[i] = s;
  }
  for (int i = 0; i < S - 1; ++i) {
    for (int j = i + 1; j < S; ++j) {
      double sxy = 0.0;
      for (int m = 0; m < n; ++m) sxy = sxy + x2[static_cast<size_t>(i) * n + m] * x2[static_cast<size_t>(j) * n + m];
      if (sxy < 0.0) {
        for (int s = 0; s < S; ++s) amp[s] = amp_stdv[s] = 0.0;
        return;
      }
      rel[static_cast<size_t>(i) * S + j] = std::log(sxy / sxx[i]);
      rel[static_cast<size_t>(j) * S + i] = -rel[static_cast<size_t>(i) * S + j];
    }
  }
  for (int i = 0; i < S; ++i) {
    double s = 0.0;
    for (int j = 0; j < S; ++j) s = s - rel[static_cast<size_t>(i) * S + j];
    amp[i] = s / S;
  }
  for (int i = 0; i < S; ++i) {
    double s = 0.0;
    for (int j = 0; j < S; ++j) {
      if (i == j) continue;
      const double d = amp[j] - amp[i] - rel[static_cast<size_t>(i) * S + j];
      s = s + d * d;
    }
    amp_stdv[i] = std::sqrt(s / (S - 2));
  }
}

// one window: x [S][n] -> t, t_stdv, amp, amp_stdv [S] (columns 4-7 of opt_data.NNNNNN.dat, src/cls_measurer.f90:388-397)
inline void measure_window(int S, int n, double dt, const double* x, double* t, double* t_stdv, double* amp,
                           double* amp_stdv, int32_t* lag_k) {
  optimize_cc(S, n, dt, x, t, t_stdv, lag_k);
  optimize_amp(S, n, dt, x, t, amp, amp_stdv);
}


// ---- scan_cc on correlation functions recomputed as hypo_tremor_correlate forms them ------------------------------
//   src/cls_correlator.f90:200-233   per window and station: taper, remove the mean, divide by the Euclidean length;
//                                    per pair: circular cross-correlation / n (FFTW there, the direct sum here), maxval
//   src/cls_correlator.f90:80        n_win = (n_smp_total - n) / n_step
//   src/cls_measurer.f90:205-236     per pair: all n * n_win values sorted, threshold = element int(n * n_win * alpha)
//                                    (1-based); a window is marked for the pair when its maximum >= threshold
//   src/cls_measurer.f90:247-253     detected when more than n_pair_thred pairs are marked
// env [S][n_total]; cc_thred [P]; cc_max [P][n_win]; detected, n_above [n_win]
inline void detect_windows(int S, long n_total, const double* env, int n, int n_step, double alpha, int n_pair_thred,
                           int n_win, double* cc_thred, double* cc_max, int32_t* detected, int32_t* n_above) {
  const int P = S * (S - 1) / 2;
  const size_t N = static_cast<size_t>(n) * n_win;
  std::vector<double> a(static_cast<size_t>(S) * n), tmp(n);
  std::vector<std::vector<double>> histo(P, std::vector<double>(N));
  for (int w = 0; w < n_win; ++w) {
    const long j1 = static_cast<long>(w) * n_step;
    for (int i = 0; i < S; ++i) {
      apply_taper(n, env + static_cast<size_t>(i) * n_total + j1, tmp.data());
      double sum = 0.0, l = 0.0;
      for (int m = 0; m < n; ++m) sum = sum + tmp[m];
      for (int m = 0; m < n; ++m) tmp[m] = tmp[m] - sum / n;
      for (int m = 0; m < n; ++m) l = l + tmp[m] * tmp[m];
      l = std::sqrt(l);
      for (int m = 0; m < n; ++m) a[static_cast<size_t>(i) * n + m] = l != 0.0 ? tmp[m] / l : 0.0;
    }
    for (int i = 0; i < S - 1; ++i) {
      for (int j = i + 1; j < S; ++j) {
        const int p = pair_index(S, i, j);
        const double *ai = a.data() + static_cast<size_t>(i) * n, *aj = a.data() + static_cast<size_t>(j) * n;
        double mx = 0.0;
        for (int k = 0; k < n; ++k) {
          double s = 0.0;
          for (int m = 0; m < n - k; ++m) s = s + ai[m] * aj[m + k];
          for (int m = n - k; m < n; ++m) s = s + ai[m] * aj[m + k - n];
          histo[p][static_cast<size_t>(w) * n + k] = s;
          if (k == 0 || s > mx) mx = s;
        }
        cc_max[static_cast<size_t>(p) * n_win + w] = mx;
      }
    }
  }
  const long rank1 = static_cast<long>(static_cast<double>(N) * alpha);
  for (int p = 0; p < P; ++p) {
    std::sort(histo[p].begin(), histo[p].end());
    cc_thred[p] = histo[p][rank1 - 1];
  }
  for (int w = 0; w < n_win; ++w) {
    int c = 0;
    for (int p = 0; p < P; ++p) c += cc_max[static_cast<size_t>(p) * n_win + w] >= cc_thred[p] ? 1 : 0;
    n_above[w] = c;
    detected[w] = c > n_pair_thred ? 1 : 0;
  }
}

}  // namespace hto
