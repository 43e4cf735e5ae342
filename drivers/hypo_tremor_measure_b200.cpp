// hypo_tremor_measure_b200 -- file-level twin of the reference's hypo_tremor_measure (src/hypo_tremor_measure.f90:25-64)
// on libhtm_b200: detection (scan_cc) and lag / amplitude optimisation (measure_lag_time) of all time windows, from the
// merged envelopes alone.
//
//   hypo_tremor_measure_b200 <parameter file> [--little-endian] [--device N] [--dry-run]
//
// Inputs in the working directory: the parameter file's station_file and one STA.merged.env per station
// (hypo_tremor_convert's output).  The reference additionally reads STA1.STA2.corr and STA1.STA2.max_corr of every
// station pair, which hypo_tremor_correlate must have written first; htm_detect_windows recomputes those correlation
// functions on the device, so that program and its files are not needed.  Outputs, as the reference
// (src/cls_measurer.f90:285-304, 386-397): detected_win.dat, cc_thred.dat, opt_data.NNNNNN.dat per detected window.
// The per-window plot files trace.NNNNNN.dat (:337-384) are not written.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../include/htm_b200.h"
#include "htm_files.hpp"

static void check(int32_t rc, const char* where) {
  if (rc == HTM_OK) return;
  char buf[512];
  htm_last_error(nullptr, buf, sizeof(buf));
  std::fprintf(stderr, "ERROR: %s: %s\n", where, buf);
  std::exit(1);
}

int main(int argc, char** argv) {
  std::string param_file;
  bool big_endian = true, dry = false;
  int device = 0;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--little-endian") big_endian = false;
    else if (a == "--dry-run") dry = true;
    else if (a == "--device" && i + 1 < argc) device = std::atoi(argv[++i]);
    else if (param_file.empty()) param_file = a;
    else param_file = "?";
  }
  if (param_file.empty() || param_file == "?") {
    std::fprintf(stderr, "USAGE: hypo_tremor_mcmc [parameter file]\n");  // the reference's own text (:27)
    return 2;
  }
  try {
    htmio::ParamFile para;
    para.read(param_file, {"n_procs", "station_file", "t_win_corr", "t_step_corr", "alpha", "n_pair_thred"});
    htmio::Stations sta;
    sta.read(para.str("station_file"));
    const int S = static_cast<int>(sta.name.size());
    if (S < 3) throw std::runtime_error("at least three stations are needed");
    // check_files (src/cls_measurer.f90:117-151): every envelope exists, one sampling interval
    std::vector<double> env, one;
    double dt = 0.0;
    size_t n_total = 0;
    for (int i = 0; i < S; ++i) {
      double dti = 0.0;
      htmio::read_envelope(sta.name[i] + ".merged.env", big_endian, one, &dti);
      if (i == 0) {
        dt = dti;
        n_total = one.size();
        env.resize(static_cast<size_t>(S) * n_total);
      } else {
        if (std::fabs(dti - dt) > 1.0e-8) throw std::runtime_error("invalid delta in envelope file");
        if (one.size() != n_total) throw std::runtime_error("invalid number of samples in " + sta.name[i] + ".merged.env");
      }
      std::copy(one.begin(), one.end(), env.begin() + static_cast<size_t>(i) * n_total);
    }
    const double t_win = para.real("t_win_corr"), t_step = para.real("t_step_corr"), alpha = para.real("alpha");
    const int n_pair_thred = para.integer("n_pair_thred");
    const int n = static_cast<int>(std::lround(t_win / dt)), n_step = static_cast<int>(std::lround(t_step / dt));
    if (n < 2 || n_step < 1 || static_cast<size_t>(n) > n_total) throw std::runtime_error("window longer than the data");
    const int n_win = static_cast<int>((n_total - n) / n_step);  // src/cls_correlator.f90:80
    const size_t P = static_cast<size_t>(S) * (S - 1) / 2;
    if (dry) {
      std::printf("{\"n_sta\": %d, \"n_total\": %zu, \"dt\": %.17g, \"n\": %d, \"n_step\": %d, \"n_win\": %d, \"n_pair\": %zu, "
                  "\"alpha\": %.17g, \"n_pair_thred\": %d, \"env_first\": %.17g, \"env_last\": %.17g}\n",
                  S, n_total, dt, n, n_step, n_win, P, alpha, n_pair_thred, env.front(), env.back());
      return 0;
    }
    if (n_win < 1) throw std::runtime_error("no complete window in the data");
    // ---- scan_cc ----
    std::vector<double> thr(P);
    std::vector<int32_t> det(n_win);
    double ms_detect = 0.0, ms_measure = 0.0;
    check(htm_detect_windows(device, S, static_cast<int64_t>(n_total), env.data(), n, n_step, alpha, n_pair_thred, n_win,
                             thr.data(), nullptr, det.data(), nullptr, &ms_detect), "htm_detect_windows");
    std::vector<int32_t> win_id;
    for (int w = 0; w < n_win; ++w)
      if (det[w]) win_id.push_back(w + 1);
    std::printf(" # of detected events: %zu out of %d\n", win_id.size(), n_win);
    {
      FILE* f = std::fopen("detected_win.dat", "w");
      if (!f) throw std::runtime_error("cannot create detected_win.dat");
      for (int32_t id : win_id) std::fprintf(f, "%12d %25.16E\n", id, (id - 1) * t_step + 0.5 * t_win);
      std::fclose(f);
      f = std::fopen("cc_thred.dat", "w");
      if (!f) throw std::runtime_error("cannot create cc_thred.dat");
      size_t p = 0;
      for (int i = 0; i < S - 1; ++i)
        for (int j = i + 1; j < S; ++j, ++p) std::fprintf(f, " %s   %s %25.16E\n", sta.name[i].c_str(), sta.name[j].c_str(), thr[p]);
      std::fclose(f);
    }
    // ---- measure_lag_time ----
    if (!win_id.empty()) {
      const size_t W = win_id.size();
      std::vector<double> t(W * S), ts(W * S), am(W * S), as(W * S);
      check(htm_measure_windows(device, S, static_cast<int64_t>(n_total), env.data(), dt, n, n_step, static_cast<int32_t>(W),
                                win_id.data(), t.data(), ts.data(), am.data(), as.data(), nullptr, &ms_measure),
            "htm_measure_windows");
      for (size_t w = 0; w < W; ++w) {
        char name[64];
        std::snprintf(name, sizeof(name), "opt_data.%06d.dat", win_id[w]);
        FILE* f = std::fopen(name, "w");
        if (!f) throw std::runtime_error(std::string("cannot create ") + name);
        for (int i = 0; i < S; ++i) {
          const size_t o = w * S + i;
          std::fprintf(f, "%25.16E %25.16E %25.16E %25.16E %25.16E %25.16E %25.16E\n", sta.x[i], sta.y[i], sta.z[i], t[o], ts[o],
                       am[o], as[o]);
        }
        std::fclose(f);
      }
    }
    std::fprintf(stderr, "hypo_tremor_measure_b200: %d windows x %zu pairs x %d lags: detection %.2f ms, measurement of %zu windows "
                         "%.2f ms on the device\n", n_win, P, n, ms_detect, win_id.size(), ms_measure);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "ERROR: %s\n", e.what());
    return 1;
  }
  return 0;
}
