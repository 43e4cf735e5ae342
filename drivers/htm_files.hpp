// File formats around the hypo_tremor_mcmc hot path, restated for the C++ twin of the driver
// (the Fortran driver keeps using the reference's own cls_param / cls_obs_data):
//   parameter file   `name = value`, `#` comments, ALL blanks removed before parsing
//                    (src/cls_line_text.f90:88-148); list-directed values (`200.d0`, `250`, `T`);
//                    unknown names are fatal (src/cls_param.f90:529-535); 29 required keys for
//                    the mcmc program (src/cls_param.f90:127-137)
//   station file     one station per line: name x y z amp_fac(1:2)   (src/cls_param.f90:350-390)
//   selected_win.dat `id  time` per line                              (src/hypo_tremor_mcmc.f90:75-87)
//   opt_data.NNNNNN.dat  n_sta lines: X Y Z t t_stdv a a_stdv          (src/cls_obs_data.f90:83-112)
//   *.RR.out         stream records int32 iter + float64[...]          (src/hypo_tremor_mcmc.f90:216-233,270-280)
//   proposal_count.txt  '(A,2I20)' of '"label"', n_propose, n_accept   (src/cls_parallel.f90:270-278 uses 2I10: the
//                       batched modes count every event and cold chain, so 10 digits overflow -- both drivers widen it)
#pragma once
#include <algorithm>
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <mutex>
#include <thread>
#include <vector>

namespace htmio {

inline std::string fortran_real_token(std::string v) {
  // list-directed input accepts a D exponent: 200.d0, 1.5D-3
  for (char& c : v)
    if (c == 'd' || c == 'D') c = 'e';
  return v;
}
inline double parse_real(const std::string& name, const std::string& v) {
  const std::string t = fortran_real_token(v);
  char* end = nullptr;
  const double x = std::strtod(t.c_str(), &end);
  if (end == t.c_str() || *end != '\0') throw std::runtime_error("bad real for " + name + ": " + v);
  return x;
}
inline int parse_int(const std::string& name, const std::string& v) {
  char* end = nullptr;
  const long x = std::strtol(v.c_str(), &end, 10);
  if (end == v.c_str() || *end != '\0') throw std::runtime_error("bad integer for " + name + ": " + v);
  return static_cast<int>(x);
}
inline bool parse_logical(const std::string& name, const std::string& v) {
  // list-directed logical: optional '.', then T/t or F/f, rest ignored (T, F, .true., .FALSE.)
  size_t i = 0;
  if (i < v.size() && v[i] == '.') ++i;
  if (i < v.size() && (v[i] == 'T' || v[i] == 't')) return true;
  if (i < v.size() && (v[i] == 'F' || v[i] == 'f')) return false;
  throw std::runtime_error("bad logical for " + name + ": " + v);
}

struct ParamFile {
  std::map<std::string, std::string> kv;

  static const std::set<std::string>& known() {
    static const std::set<std::string> k = {
        "station_file", "time_id_file", "cmp1", "cmp2", "data_dir", "filename_format", "n_procs", "t_win_conv",
        "t_win_corr", "t_step_corr", "n_pair_thred", "alpha", "vs_min", "vs_max", "b_min", "b_max", "z_guess",
        "n_iter", "n_burn", "n_interval", "n_chains", "n_cool", "temp_high", "prior_width_xy", "prior_width_z",
        "prior_z", "prior_vs", "prior_width_vs", "prior_qs", "prior_width_qs", "prior_t_corr", "prior_width_t_corr",
        "prior_a_corr", "prior_width_a_corr", "step_size_xy", "step_size_z", "step_size_vs", "step_size_t_corr",
        "step_size_qs", "step_size_a_corr", "solve_vs", "solve_qs", "solve_t_corr", "solve_a_corr", "use_amp",
        "use_time"};
    return k;
  }
  static const std::vector<std::string>& required_mcmc() {
    static const std::vector<std::string> r = {
        "n_procs", "station_file", "n_iter", "n_burn", "n_interval", "n_chains", "n_cool", "temp_high", "prior_z",
        "prior_width_z", "prior_width_xy", "prior_vs", "prior_width_vs", "prior_qs", "prior_width_qs",
        "prior_width_t_corr", "prior_width_a_corr", "step_size_z", "step_size_xy", "step_size_vs", "step_size_qs",
        "step_size_t_corr", "step_size_a_corr", "solve_vs", "solve_t_corr", "solve_qs", "solve_a_corr", "use_time",
        "use_amp"};
    return r;
  }

  void read(const std::string& path) { read(path, required_mcmc()); }
  // required: the keys the calling program needs (src/cls_param.f90 checks per from_where)
  void read(const std::string& path, const std::vector<std::string>& required) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("cannot open parameter file " + path);
    std::string line;
    while (std::getline(in, line)) {
      const size_t hash = line.find('#');
      if (hash != std::string::npos) line.erase(hash);
      line.erase(std::remove(line.begin(), line.end(), ' '), line.end());   // every blank, not just the ends
      line.erase(std::remove(line.begin(), line.end(), '\r'), line.end());
      line.erase(std::remove(line.begin(), line.end(), '\t'), line.end());
      if (line.empty()) continue;
      const size_t eq = line.find('=');
      if (eq == std::string::npos || eq == 0 || eq + 1 == line.size()) continue;  // is_ok = .false.: ignored
      const std::string name = line.substr(0, eq), val = line.substr(eq + 1);
      if (!known().count(name)) throw std::runtime_error("Invalid parameter name : " + name + "  (?)");
      kv[name] = val;
    }
    for (const std::string& k : required)
      if (!kv.count(k)) throw std::runtime_error(k + " is not given.");
  }
  bool has(const std::string& k) const { return kv.count(k) != 0; }
  const std::string& str(const std::string& k) const { return kv.at(k); }
  double real(const std::string& k) const { return parse_real(k, kv.at(k)); }
  double real_or(const std::string& k, double d) const { return has(k) ? real(k) : d; }  // prior_t_corr / prior_a_corr
  int integer(const std::string& k) const { return parse_int(k, kv.at(k)); }
  bool logical(const std::string& k) const { return parse_logical(k, kv.at(k)); }
};

struct Stations {
  std::vector<std::string> name;
  std::vector<double> x, y, z;
  void read(const std::string& path) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("cannot open " + path);
    std::string line;
    while (std::getline(in, line)) {
      std::istringstream ss(line);
      std::string nm, sx, sy, sz;
      if (!(ss >> nm >> sx >> sy >> sz)) {
        if (line.find_first_not_of(" \t\r") == std::string::npos) continue;
        throw std::runtime_error("bad station line: " + line);
      }
      name.push_back(nm);
      x.push_back(parse_real("sta_x", sx));
      y.push_back(parse_real("sta_y", sy));
      z.push_back(parse_real("sta_z", sz));
    }
  }
};

inline std::vector<int> read_selected_windows(const std::string& path) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("cannot open " + path);
  std::vector<int> ids;
  std::string a, b;
  while (in >> a >> b) ids.push_back(parse_int("win_id", a));
  return ids;
}

// detected_win.dat / selected_win.dat: window number and time, list-directed (src/cls_measurer.f90:289-292)
inline void read_window_list(const std::string& path, std::vector<int>& ids, std::vector<double>& times) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("cannot open " + path);
  std::string a, b;
  while (in >> a >> b) {
    ids.push_back(parse_int("win_id", a));
    times.push_back(parse_real("win_time", b));
  }
}

// STA.merged.env: unformatted stream of (time, envelope) float64 pairs (src/cls_convertor.f90:259-272), read by the
// reference with direct access, recl = 8 (src/cls_measurer.f90:132-139, 345-349)
inline void read_envelope(const std::string& path, bool big_endian, std::vector<double>& v, double* dt) {
  std::ifstream in(path, std::ios::binary | std::ios::ate);
  if (!in) throw std::runtime_error(path + " does not exist");
  const std::streamsize bytes = in.tellg();
  if (bytes < 32 || bytes % 16 != 0) throw std::runtime_error("bad envelope file " + path);
  std::vector<unsigned char> raw(static_cast<size_t>(bytes));
  in.seekg(0);
  in.read(reinterpret_cast<char*>(raw.data()), bytes);
  auto at = [&](size_t k) {
    unsigned char b[8];
    for (int i = 0; i < 8; ++i) b[i] = raw[8 * k + (big_endian ? 7 - i : i)];
    double d;
    std::memcpy(&d, b, 8);
    return d;
  };
  const size_t n = static_cast<size_t>(bytes) / 16;
  v.resize(n);
  for (size_t i = 0; i < n; ++i) v[i] = at(2 * i + 1);
  *dt = at(2) - at(0);
}

// obs arrays in Fortran (n_sta, n_events) column-major order == C [n_events][n_sta]
struct Observations {
  std::vector<double> t_obs, t_stdv, a_obs, a_stdv;
  // One opt_data.NNNNNN.dat per event (src/cls_obs_data.f90:83-112 opens them one after the other on every
  // rank).  Here the events are split over host threads and every file is read in one piece: at 10^4-10^5
  // events this stage would otherwise dominate the wall clock once the chains run on the GPU.
  void read(const std::vector<int>& win_id, int n_sta, const std::string& dir = ".", unsigned n_threads = 0) {
    const size_t E = win_id.size();
    t_obs.resize(E * n_sta);
    t_stdv.resize(E * n_sta);
    a_obs.resize(E * n_sta);
    a_stdv.resize(E * n_sta);
    if (n_threads == 0) n_threads = std::max(1u, std::thread::hardware_concurrency());
    n_threads = static_cast<unsigned>(std::min<size_t>(n_threads, std::max<size_t>(1, E / 64)));
    std::mutex err_lock;
    std::string first_error;
    size_t first_error_event = E;
    auto work = [&](size_t i0, size_t i1) {
      std::string buf, tok;
      char fname[64];
      for (size_t i = i0; i < i1; ++i) {
        try {
          std::snprintf(fname, sizeof(fname), "opt_data.%06d.dat", win_id[i]);
          const std::string path = dir + "/" + fname;
          FILE* f = std::fopen(path.c_str(), "rb");
          if (!f) throw std::runtime_error(std::string("ERROR: obs_file is not found: ") + fname);
          buf.clear();
          char chunk[16384];
          size_t got;
          while ((got = std::fread(chunk, 1, sizeof(chunk), f)) > 0) buf.append(chunk, got);
          std::fclose(f);
          // one list-directed read per station record, as `read(io,*)` of 7 items does (src/cls_obs_data.f90:92-99):
          // the 7 items may continue over following lines; whatever follows the 7th on its line is skipped, and
          // the next station starts on a new line
          size_t pos = 0;
          auto next = [&]() -> bool {
            while (pos < buf.size() && (std::isspace(static_cast<unsigned char>(buf[pos])) || buf[pos] == ',')) ++pos;
            if (pos >= buf.size()) return false;
            const size_t b0 = pos;
            while (pos < buf.size() && !std::isspace(static_cast<unsigned char>(buf[pos])) && buf[pos] != ',') ++pos;
            tok.assign(buf, b0, pos - b0);
            return true;
          };
          for (int j = 0; j < n_sta; ++j) {
            const size_t k = i * n_sta + j;
            int c = 0;
            while (c < 7) {
              if (!next()) throw std::runtime_error(std::string("short obs file ") + fname);
              // r*c repeat form of list-directed input (e.g. 2*0.0)
              int rep = 1;
              std::string val = tok;
              const size_t star = tok.find('*');
              if (star != std::string::npos && star > 0 && tok.find_first_not_of("0123456789") == star) {
                rep = std::atoi(tok.substr(0, star).c_str());
                val = tok.substr(star + 1);
              }
              for (int q = 0; q < rep && c < 7; ++q, ++c) {
                if (c == 3) t_obs[k] = parse_real("t_obs", val);
                if (c == 4) t_stdv[k] = parse_real("t_stdv", val);
                if (c == 5) a_obs[k] = parse_real("a_obs", val);
                if (c == 6) a_stdv[k] = parse_real("a_stdv", val);
              }
            }
            while (pos < buf.size() && buf[pos] != '\n') ++pos;  // rest of the record's last line
          }
        } catch (const std::exception& e) {
          std::lock_guard<std::mutex> g(err_lock);
          if (i < first_error_event) {  // report what the sequential reader would have hit first
            first_error_event = i;
            first_error = e.what();
          }
          return;
        }
      }
    };
    if (n_threads <= 1) {
      work(0, E);
    } else {
      std::vector<std::thread> pool;
      for (unsigned t = 0; t < n_threads; ++t) pool.emplace_back(work, E * t / n_threads, E * (t + 1) / n_threads);
      for (auto& th : pool) th.join();
    }
    if (first_error_event < E) throw std::runtime_error(first_error);
  }
  // obs%make_initial_guess (src/cls_obs_data.f90:120-134): station of the FIRST maximum of a_obs
  void initial_guess(const Stations& st, std::vector<double>& x_mu, std::vector<double>& y_mu) const {
    const size_t S = st.x.size(), E = a_obs.size() / S;
    x_mu.resize(E);
    y_mu.resize(E);
    for (size_t i = 0; i < E; ++i) {
      size_t best = 0;
      for (size_t j = 1; j < S; ++j)
        if (a_obs[i * S + j] > a_obs[i * S + best]) best = j;
      x_mu[i] = st.x[best];
      y_mu[i] = st.y[best];
    }
  }
};

// stream-access unformatted records; big-endian when the reference is built with its shipped
// flags (-fconvert=big-endian, src/Makefile:9,14,18)
struct StreamFile {
  FILE* f = nullptr;
  bool big_endian = true;
  void open(const std::string& path, bool be) {
    f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot create " + path);
    big_endian = be;
  }
  void close() {
    if (f) std::fclose(f);
    f = nullptr;
  }
  template <typename T>
  void put(T v) {
    unsigned char b[sizeof(T)];
    std::memcpy(b, &v, sizeof(T));
    if (big_endian) std::reverse(b, b + sizeof(T));
    std::fwrite(b, 1, sizeof(T), f);
  }
  void record(int32_t iter, const double* v, size_t n) {
    put<int32_t>(iter);
    for (size_t i = 0; i < n; ++i) put<double>(v[i]);
  }
};

}  // namespace htmio
