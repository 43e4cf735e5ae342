// hypo_tremor_select_b200 -- file-level twin of the reference's hypo_tremor_select (src/hypo_tremor_select.f90) on
// libhtm_b200: the regression / acceptance stage over all detected windows in one call.
//
//   hypo_tremor_select_b200 <parameter file> [--device N] [--loader-threads N]
//
// Inputs in the working directory: the parameter file's station_file, detected_win.dat, opt_data.NNNNNN.dat per detected
// window (hypo_tremor_measure's output).  Outputs (:107-130): regress.dat (window, vs, B, t0, a0, cc_t, cc_a) and
// selected_win.dat (window, time) for the windows with vs_min <= vs <= vs_max and b_min <= B <= b_max.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../include/htm_b200.h"
#include "htm_files.hpp"

int main(int argc, char** argv) {
  std::string param_file;
  int device = 0;
  unsigned loader_threads = 0;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--device" && i + 1 < argc) device = std::atoi(argv[++i]);
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
    para.read(param_file, {"n_procs", "station_file", "z_guess", "vs_min", "vs_max", "b_min", "b_max"});
    htmio::Stations sta;
    sta.read(para.str("station_file"));
    const int S = static_cast<int>(sta.name.size());
    std::vector<int> win_id;
    std::vector<double> win_t;
    htmio::read_window_list("detected_win.dat", win_id, win_t);
    const int E = static_cast<int>(win_id.size());
    FILE* fr = std::fopen("regress.dat", "w");
    FILE* fs = std::fopen("selected_win.dat", "w");
    if (!fr || !fs) throw std::runtime_error("cannot create regress.dat / selected_win.dat");
    if (E > 0) {
      htmio::Observations obs;
      obs.read(win_id, S, ".", loader_threads);
      std::vector<double> vs(E), t0(E), b(E), a0(E), cct(E), cca(E);
      std::vector<int32_t> sel(E);
      double ms = 0.0;
      const int32_t rc = htm_select_events(device, S, E, sta.x.data(), sta.y.data(), sta.z.data(), para.real("z_guess"),
                                           obs.t_obs.data(), obs.t_stdv.data(), obs.a_obs.data(), obs.a_stdv.data(),
                                           para.real("vs_min"), para.real("vs_max"), para.real("b_min"), para.real("b_max"),
                                           vs.data(), t0.data(), b.data(), a0.data(), cct.data(), cca.data(), sel.data(), &ms);
      if (rc != HTM_OK) {
        char buf[512];
        htm_last_error(nullptr, buf, sizeof(buf));
        throw std::runtime_error(std::string("htm_select_events: ") + buf);
      }
      for (int i = 0; i < E; ++i) {
        std::fprintf(fr, "%12d %25.16E %25.16E %25.16E %25.16E %25.16E %25.16E\n", win_id[i], vs[i], b[i], t0[i], a0[i], cct[i],
                     cca[i]);
        if (sel[i]) std::fprintf(fs, "%12d %25.16E\n", win_id[i], win_t[i]);
      }
      std::fprintf(stderr, "hypo_tremor_select_b200: %d windows x %d stations: %.3f ms on the device\n", E, S, ms);
    }
    std::fclose(fr);
    std::fclose(fs);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "ERROR: %s\n", e.what());
    return 1;
  }
  return 0;
}
