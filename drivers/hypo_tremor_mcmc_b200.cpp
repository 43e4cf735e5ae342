// hypo_tremor_mcmc_b200 -- C++ twin of the Fortran drop-in driver
// (hypotremormcmc_b200/fortran/hypo_tremor_mcmc_b200.f90).  It makes the SAME C-ABI calls in the
// same order, so the boundary is exercised end to end in an image that has no Fortran compiler.
//
//   hypo_tremor_mcmc_b200 <parameter file> [--precision 32|64] [--seed N] [--little-endian]
//                         [--chunk RECORDS] [--dry-run] [--loader-threads N] [--summary]
//
// --summary additionally writes hypo.stat, station_corrections.stat and uniform_structure.stat -- the tables
// hypo_tremor_statistics computes from the .out files (src/cls_statistics.f90:216-264,345-431) -- from the
// device-side posterior store (htm_posterior_quantiles), without re-reading the samples.
//
// Inputs in the working directory as for the reference (src/hypo_tremor_mcmc.f90:53-98): the
// parameter file's station_file, selected_win.dat, opt_data.NNNNNN.dat.  Outputs: hypo.RR.out,
// t_corr.RR.out, vs.RR.out, a_corr.RR.out, qs.RR.out, likelihoodRR.out per virtual rank RR, and
// proposal_count.txt.  n_procs of the parameter file = number of virtual ranks (no MPI).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../include/htm_b200.h"
#include "htm_files.hpp"

static void check(htm_handle h, int32_t rc, const char* where) {
  if (rc == HTM_OK) return;
  char buf[512];
  htm_last_error(h, buf, sizeof(buf));
  std::fprintf(stderr, "ERROR: %s: %s\n", where, buf);   // print and stop, like the reference
  std::exit(1);
}

int main(int argc, char** argv) {
  std::string param_file;
  int precision = 32, chunk = 0;  // 0 = choose from the record size
  unsigned long long seed = 20231001ull;
  bool big_endian = true, dry = false, summary = false;
  unsigned loader_threads = 0;  // 0 = all host cores
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--precision" && i + 1 < argc) precision = std::atoi(argv[++i]);
    else if (a == "--seed" && i + 1 < argc) seed = std::strtoull(argv[++i], nullptr, 10);
    else if (a == "--chunk" && i + 1 < argc) chunk = std::atoi(argv[++i]);
    else if (a == "--little-endian") big_endian = false;
    else if (a == "--dry-run") dry = true;
    else if (a == "--summary") summary = true;
    else if (a == "--loader-threads" && i + 1 < argc) loader_threads = static_cast<unsigned>(std::atoi(argv[++i]));
    else if (param_file.empty()) param_file = a;
    else param_file = "?";
  }
  if (param_file.empty() || param_file == "?") {
    std::fprintf(stderr, "USAGE: hypo_tremor_mcmc [parameter file]\n");
    return 2;
  }
  try {
    htmio::ParamFile para;
    para.read(param_file);
    htmio::Stations sta;
    sta.read(para.str("station_file"));
    const std::vector<int> win_id = htmio::read_selected_windows("selected_win.dat");
    const int n_sta = static_cast<int>(sta.x.size()), n_events = static_cast<int>(win_id.size());
    htmio::Observations obs;
    const auto t_load = std::chrono::steady_clock::now();
    obs.read(win_id, n_sta, ".", loader_threads);
    std::fprintf(stderr, "read %d observation files in %.3f s\n", n_events,
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t_load).count());
    std::vector<double> x_mu, y_mu;
    obs.initial_guess(sta, x_mu, y_mu);

    htm_config cfg;
    htm_config_default(&cfg);
    cfg.seed = seed;
    cfg.n_sta = n_sta;
    cfg.n_events = n_events;
    cfg.n_procs = para.integer("n_procs");
    cfg.n_chains = para.integer("n_chains");
    cfg.n_cool = para.integer("n_cool");
    cfg.temp_high = para.real("temp_high");
    cfg.n_iter = para.integer("n_iter");
    cfg.n_burn = para.integer("n_burn");
    cfg.n_interval = para.integer("n_interval");
    cfg.prior_z = para.real("prior_z");
    cfg.prior_width_z = para.real("prior_width_z");
    cfg.prior_width_xy = para.real("prior_width_xy");
    cfg.prior_vs = para.real("prior_vs");
    cfg.prior_width_vs = para.real("prior_width_vs");
    cfg.prior_qs = para.real("prior_qs");
    cfg.prior_width_qs = para.real("prior_width_qs");
    cfg.prior_t_corr = para.real_or("prior_t_corr", 0.0);   // optional, default 0 (src/cls_param.f90:89,91)
    cfg.prior_width_t_corr = para.real("prior_width_t_corr");
    cfg.prior_a_corr = para.real_or("prior_a_corr", 0.0);
    cfg.prior_width_a_corr = para.real("prior_width_a_corr");
    cfg.step_size_z = para.real("step_size_z");
    cfg.step_size_xy = para.real("step_size_xy");
    cfg.step_size_vs = para.real("step_size_vs");
    cfg.step_size_qs = para.real("step_size_qs");
    cfg.step_size_t_corr = para.real("step_size_t_corr");
    cfg.step_size_a_corr = para.real("step_size_a_corr");
    cfg.solve_vs = para.logical("solve_vs");
    cfg.solve_t_corr = para.logical("solve_t_corr");
    cfg.solve_qs = para.logical("solve_qs");
    cfg.solve_a_corr = para.logical("solve_a_corr");
    cfg.use_time = para.logical("use_time");
    cfg.use_amp = para.logical("use_amp");
    const bool any_solve = cfg.solve_vs || cfg.solve_t_corr || cfg.solve_qs || cfg.solve_a_corr;
    cfg.mode = any_solve ? HTM_MODE_BLOCKED_GIBBS : HTM_MODE_FACTORISED;
    cfg.precision = precision;
    if (chunk <= 0) {
      // records drained per htm_run: 64, fewer when one hypo record (3 E doubles per cold chain) is large --
      // the host buffers below stay within about 256 MB
      const double per_rec = 24.0 * n_events * cfg.n_cool * cfg.n_procs;
      chunk = static_cast<int>(std::max(1.0, std::min(64.0, 2.56e8 / per_rec - 1.0)));
    }
    cfg.max_samples = chunk + 1;
    cfg.summary = summary ? 1 : 0;

    if (dry) {  // parse-only: what the driver understood, as JSON (used by the CPU tests)
      std::printf("{\"n_sta\": %d, \"n_events\": %d, \"n_procs\": %d, \"n_chains\": %d, \"n_cool\": %d, "
                  "\"n_iter\": %d, \"n_burn\": %d, \"n_interval\": %d, \"temp_high\": %.17g, \"prior_vs\": %.17g, "
                  "\"prior_qs\": %.17g, \"prior_t_corr\": %.17g, \"step_size_xy\": %.17g, \"solve_vs\": %d, "
                  "\"use_amp\": %d, \"mode\": %d, \"x_mu0\": %.17g, \"y_mu0\": %.17g, \"t_obs00\": %.17g, "
                  "\"a_stdv_last\": %.17g, \"sta_z_last\": %.17g}\n",
                  n_sta, n_events, cfg.n_procs, cfg.n_chains, cfg.n_cool, cfg.n_iter, cfg.n_burn, cfg.n_interval,
                  cfg.temp_high, cfg.prior_vs, cfg.prior_qs, cfg.prior_t_corr, cfg.step_size_xy, cfg.solve_vs,
                  cfg.use_amp, cfg.mode, x_mu[0], y_mu[0], obs.t_obs[0], obs.a_stdv.back(), sta.z.back());
      return 0;
    }

    htm_handle h = nullptr;
    check(nullptr, htm_create(&h, &cfg), "htm_create");
    check(h, htm_set_stations(h, sta.x.data(), sta.y.data(), sta.z.data()), "htm_set_stations");
    check(h, htm_set_observations(h, obs.t_obs.data(), obs.t_stdv.data(), obs.a_obs.data(), obs.a_stdv.data()),
          "htm_set_observations");
    check(h, htm_set_xy_prior(h, x_mu.data(), y_mu.data()), "htm_set_xy_prior");
    check(h, htm_init_chains(h), "htm_init_chains");

    const int R = cfg.n_procs;
    std::vector<htmio::StreamFile> f_hypo(R), f_tc(R), f_vs(R), f_ac(R), f_qs(R), f_lik(R);
    char nm[64];
    for (int r = 0; r < R; ++r) {
      auto open = [&](htmio::StreamFile& f, const char* stem) {
        std::snprintf(nm, sizeof(nm), "%s%02d.out", stem, r);
        f.open(nm, big_endian);
      };
      open(f_hypo[r], "hypo.");
      open(f_tc[r], "t_corr.");
      open(f_vs[r], "vs.");
      open(f_ac[r], "a_corr.");
      open(f_qs[r], "qs.");
      open(f_lik[r], "likelihood");   // sic: no dot before the rank (reference file name)
    }
    std::printf(" start MCMC\n");
    const int cap = (chunk + 1) * cfg.n_cool * R;
    std::vector<int32_t> it(cap);
    std::vector<double> vs(cap), qs(cap), lk(cap), hy(static_cast<size_t>(cap) * 3 * n_events),
        tc(static_cast<size_t>(cap) * n_sta), ac(static_cast<size_t>(cap) * n_sta);
    const int n_int = cfg.n_interval > 1 ? cfg.n_interval : 1;
    for (int it0 = 1; it0 <= cfg.n_iter;) {
      const long last = static_cast<long>(it0) + static_cast<long>(chunk) * n_int - 1;
      const int it1 = last < cfg.n_iter ? static_cast<int>(last) : cfg.n_iter;
      check(h, htm_run(h, it0, it1), "htm_run");
      for (int r = 0; r < R; ++r) {
        int32_t n = 0;
        check(h, htm_fetch_samples(h, r, cap, &n, it.data(), vs.data(), qs.data(), hy.data(), tc.data(), ac.data()),
              "htm_fetch_samples");
        for (int j = 0; j < n; ++j) {   // the reference's write order, src/hypo_tremor_mcmc.f90:273-277
          f_vs[r].record(it[j], &vs[j], 1);
          f_hypo[r].record(it[j], &hy[static_cast<size_t>(j) * 3 * n_events], 3 * static_cast<size_t>(n_events));
          f_tc[r].record(it[j], &tc[static_cast<size_t>(j) * n_sta], n_sta);
          f_qs[r].record(it[j], &qs[j], 1);
          f_ac[r].record(it[j], &ac[static_cast<size_t>(j) * n_sta], n_sta);
        }
        check(h, htm_fetch_likelihood(h, r, cap, &n, it.data(), lk.data()), "htm_fetch_likelihood");
        for (int j = 0; j < n; ++j) f_lik[r].record(it[j], &lk[j], 1);
      }
      it0 = it1 + 1;
    }
    int64_t np[7], na[7];
    check(h, htm_get_counts(h, np, na), "htm_get_counts");
    static const char* label[7] = {"vs   ", "t_cor", "qs   ", "a_cor", "x    ", "y    ", "z    "};  // character(5)
    FILE* pc = std::fopen("proposal_count.txt", "w");
    for (int k = 0; k < 7; ++k) std::fprintf(pc, "\"%s\"%20lld%20lld\n", label[k], static_cast<long long>(np[k]), static_cast<long long>(na[k]));
    std::fclose(pc);
    if (summary) {
      int32_t n_mod = 0;
      std::vector<double> hq(static_cast<size_t>(9) * n_events), vq(3), qq(3), tq(static_cast<size_t>(3) * n_sta),
          aq(static_cast<size_t>(3) * n_sta);
      check(h, htm_posterior_quantiles(h, &n_mod, hq.data(), vq.data(), qq.data(), tq.data(), aq.data()), "htm_posterior_quantiles");
      FILE* f = std::fopen("hypo.stat", "w");   // '(I9,9F13.6)', src/cls_statistics.f90:244-255
      std::fprintf(f, "# window ID, x (50%%), x (2.5%%) x (97.5%%), y (50%%), y (2.5%%), y (97.5%%)z (50 %%), z (2.5%%), z (97.5%%)\n");
      for (int i = 0; i < n_events; ++i) {
        std::fprintf(f, "%9d", win_id[i]);
        for (int k = 0; k < 9; ++k) std::fprintf(f, "%13.6f", hq[static_cast<size_t>(9) * i + k]);
        std::fprintf(f, "\n");
      }
      std::fclose(f);
      f = std::fopen("station_corrections.stat", "w");   // '(A12,6F13.6)', :373-381
      std::fprintf(f, "# station name, t_corr (50%%), t_corr (2.5%%) t_corr (97.5%%), a_corr (50%%), a_corr (2.5%%), a_corr (97.5%%)\n");
      for (int j = 0; j < n_sta; ++j) {
        std::fprintf(f, "%12.12s", sta.name[j].c_str());
        for (int k = 0; k < 3; ++k) std::fprintf(f, "%13.6f", tq[static_cast<size_t>(3) * j + k]);
        for (int k = 0; k < 3; ++k) std::fprintf(f, "%13.6f", aq[static_cast<size_t>(3) * j + k]);
        std::fprintf(f, "\n");
      }
      std::fclose(f);
      f = std::fopen("uniform_structure.stat", "w");   // '(6F13.6)', :419-423
      std::fprintf(f, "# Vs (50%%), Vs (2.5%%) Vs (97.5%%), Qs (50%%), Qs (2.5%%), Qs (97.5%%)\n");
      for (int k = 0; k < 3; ++k) std::fprintf(f, "%13.6f", vq[k]);
      for (int k = 0; k < 3; ++k) std::fprintf(f, "%13.6f", qq[k]);
      std::fprintf(f, "\n");
      std::fclose(f);
      std::fprintf(stderr, "posterior summary of %d samples written\n", n_mod);
    }
    for (int r = 0; r < R; ++r) {
      f_hypo[r].close();
      f_tc[r].close();
      f_vs[r].close();
      f_ac[r].close();
      f_qs[r].close();
      f_lik[r].close();
    }
    htm_destroy(h);
  } catch (const std::exception& ex) {
    std::fprintf(stderr, "ERROR: %s\n", ex.what());
    return 1;
  }
  return 0;
}
