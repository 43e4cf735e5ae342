// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
//
// CPU restatement of the hypo_tremor_mcmc inversion hot path of akuhara/HypoTremorMCMC.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may use anything in oracle/.  The product (libhtm_b200.so) never links or calls it.
//
// PARITY PINNING: the reference ships no tests, fixtures or golden vectors, and it cannot
// be compiled here (no Fortran compiler, no MPI) -- so this oracle is "parity unpinned" by
// the reference itself.  It is pinned by (i) the hand-derived mod_random known answers of
// SURVEY.md section 8a, (ii) an independent numpy restatement of cls_forward
// (tests/golden/make_golden.py) and (iii) the analytic identities of SURVEY.md section 8c.
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference/).  Arithmetic follows the default -O0 build: float64, int32, no FMA
// contraction (compile with -ffp-contract=off), left-to-right sums.
#pragma once
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/htm_b200.h"

namespace hto {

// ------------------------------------------------------------------------------------
// mod_random  (src/mod_random.f90)
// ------------------------------------------------------------------------------------
struct Xorshift128 {
  // Fortran default integers are int32; ishft is a LOGICAL shift, so the state is kept as
  // uint32 and only read as signed where the reference does dble(w) (:72).
  uint32_t x = 0, y = 0, z = 0, w = 0;
  std::vector<int32_t>* tape = nullptr;  // optional recorder of every raw w produced
  uint64_t n_draws = 0;

  // init_random, src/mod_random.f90:39-55.  Seed arithmetic wraps mod 2^32.
  void init(int32_t i1, int32_t i2, int32_t i3, int32_t i4, int32_t rank) {
    const uint32_t j1 = static_cast<uint32_t>(rank + 1);
    const uint32_t p2 = j1 * j1;
    const uint32_t p4 = p2 * p2;
    auto seed = [&](int32_t i) -> uint32_t {
      const uint32_t u = static_cast<uint32_t>(i);
      return u * p4 + 1000u * u * p2 + u;
    };
    x = seed(i1);
    y = seed(i2);
    z = seed(i3);
    w = seed(i4);
    n_draws = 0;
  }

  // state update shared by rand_u and rand_u2, src/mod_random.f90:63-71 / :81-89
  inline int32_t next_raw() {
    const uint32_t t = x ^ (x << 11);
    x = y;
    y = z;
    z = w;
    w = (w ^ (w >> 19)) ^ (t ^ (t >> 8));
    ++n_draws;
    const int32_t sw = static_cast<int32_t>(w);
    if (tape) tape->push_back(sw);
    return sw;
  }
  static inline double to_u(int32_t sw) {  // :72  U[0,1)
    return (static_cast<double>(sw) + 2147483648.0) / 4294967296.0;
  }
  static inline double to_u2(int32_t sw) {  // :90  U(0,1)
    return (static_cast<double>(sw) + 2147483648.0 + 0.5) / 4294967296.0;
  }
  inline double rand_u() { return to_u(next_raw()); }
  inline double rand_u2() { return to_u2(next_raw()); }
  // rand_g, src/mod_random.f90:95-102 (v1 drawn first)
  inline double rand_g() {
    const double pi2 = 2.0 * std::acos(-1.0);
    const double v1 = rand_u2();
    const double v2 = rand_u2();
    return std::sqrt(-2.0 * std::log(v1)) * std::cos(pi2 * v2);
  }
  // rand_r, src/mod_random.f90:106-112
  inline double rand_r() {
    const double u = rand_u2();
    return std::sqrt(-2.0 * std::log(u));
  }
};

// ------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), the counter-based generator the B200 modes B/C use
// instead of mod_random.  Restated here so the oracle can run the SAME factorised /
// blocked-Gibbs schedule as the GPU on the same draws.
// ------------------------------------------------------------------------------------
struct Philox {
  static inline void round(uint32_t c[4], uint32_t k0, uint32_t k1) {
    const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c[0];
    const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c[2];
    const uint32_t hi0 = static_cast<uint32_t>(p0 >> 32), lo0 = static_cast<uint32_t>(p0);
    const uint32_t hi1 = static_cast<uint32_t>(p1 >> 32), lo1 = static_cast<uint32_t>(p1);
    const uint32_t n0 = hi1 ^ c[1] ^ k0;
    const uint32_t n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0;
    c[1] = lo1;
    c[2] = n2;
    c[3] = lo0;
  }
  static inline void gen(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                         uint32_t out[4]) {
    uint32_t c[4] = {c0, c1, c2, c3};
    uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
    for (int r = 0; r < 10; ++r) {
      round(c, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    for (int i = 0; i < 4; ++i) out[i] = c[i];
  }
  // Uniform constructions shared bit-for-bit by the f32 and f64 device kernels:
  // [0,1) with 24 bits, (0,1) with 23 bits + 1/2 (both exact in float32).
  static inline double u_co(uint32_t w) { return static_cast<double>(w >> 8) * (1.0 / 16777216.0); }
  static inline double u_oo(uint32_t w) {
    return (static_cast<double>(w >> 9) + 0.5) * (1.0 / 8388608.0);
  }
  // int(u*n) for u = u_co(w), done in integers (n < 2^24 is guaranteed by the callers
  // that use it for indices; larger n use u_co directly).
  static inline uint32_t below(uint32_t w, uint32_t n) {
    return static_cast<uint32_t>((static_cast<uint64_t>(w >> 8) * n) >> 24);
  }
};
// counter word c2 ("purpose") of the Philox streams
enum : uint32_t { PHX_STEP = 0, PHX_SWAP = 1, PHX_INIT = 2, PHX_GLOBAL = 3, PHX_TEMP = 4 };

// ------------------------------------------------------------------------------------
// cls_model  (src/cls_model.f90)
// ------------------------------------------------------------------------------------
struct Model {
  int32_t nx = 0;
  std::vector<int32_t> prior_type;  // 0 Gaussian, 1 Rayleigh-like (:9)
  std::vector<double> x, mu, sigma, step_size;

  Model() = default;
  explicit Model(int32_t n) : nx(n), prior_type(n, 0), x(n, 0.0), mu(n, 0.0), sigma(n, 0.0),
                              step_size(n, 0.0) {}
  // set_prior :65-79 (0-based i here), set_perturb :83-91
  void set_prior(int32_t i, double m, double s, int32_t type = 0) {
    mu[i] = m;
    sigma[i] = s;
    prior_type[i] = type;
  }
  // generate_model, :139-158
  void generate(Xorshift128& rng) {
    for (int32_t i = 0; i < nx; ++i) {
      if (prior_type[i] == 0) {
        x[i] = mu[i] + rng.rand_g() * sigma[i];
      } else if (prior_type[i] == 1) {
        x[i] = mu[i] + rng.rand_r() * sigma[i];
      } else {
        throw std::runtime_error("unsupported prior type");
      }
    }
  }
  // The arithmetic of model_perturb, :162-190, for a given standard normal g.
  // Returns x_new; sets log_prior_ratio and prior_ok.
  inline double perturb_with(int32_t i, double g, double& log_prior_ratio, bool& prior_ok) const {
    prior_ok = true;
    const double x_old = x[i];
    const double x_new = x_old + g * step_size[i];
    log_prior_ratio = -((x_new - mu[i]) * (x_new - mu[i]) - (x_old - mu[i]) * (x_old - mu[i])) /
                      (2.0 * sigma[i] * sigma[i]);
    if (prior_type[i] == 1) {
      if (x_new <= mu[i]) {
        log_prior_ratio = static_cast<double>(-1.0e+30f);  // single-precision literal, :180
        prior_ok = false;
      } else {
        log_prior_ratio = log_prior_ratio + std::log(x_new - mu[i]) - std::log(x_old - mu[i]);
      }
    }
    return x_new;
  }
};

// ------------------------------------------------------------------------------------
// cls_forward  (src/cls_forward.f90)
// ------------------------------------------------------------------------------------
struct Forward {
  int32_t n_sta = 0, n_events = 0;
  bool use_amp = true, use_time = true;
  std::vector<double> sta_x, sta_y, sta_z;
  // (n_sta, n_events) column-major: element (j,i) at [i*n_sta + j]
  std::vector<double> t_obs, t_stdv, t_precision, log_t_stdv;
  std::vector<double> a_obs, a_stdv, a_precision, log_a_stdv;
  mutable std::vector<double> t_syn, a_syn;  // scratch

  // log_2pi_half, src/cls_forward.f90:5
  static double log_2pi_half() { return 0.5 * std::log(2.0 * std::acos(-1.0)); }

  // init_forward, src/cls_forward.f90:40-96
  void init(int32_t ns, int32_t ne, const double* sx, const double* sy, const double* sz,
            const double* tobs, const double* tstd, const double* aobs, const double* astd,
            bool use_amp_, bool use_time_) {
    n_sta = ns;
    n_events = ne;
    use_amp = use_amp_;
    use_time = use_time_;
    sta_x.assign(sx, sx + ns);
    sta_y.assign(sy, sy + ns);
    sta_z.assign(sz, sz + ns);
    const size_t n = static_cast<size_t>(ns) * ne;
    t_obs.assign(tobs, tobs + n);
    t_stdv.assign(tstd, tstd + n);
    a_obs.assign(aobs, aobs + n);
    a_stdv.assign(astd, astd + n);
    t_precision.resize(n);
    a_precision.resize(n);
    log_t_stdv.resize(n);
    log_a_stdv.resize(n);
    for (size_t k = 0; k < n; ++k) {
      if (t_stdv[k] > 1.e-16) {  // the branch looks at t_stdv only, :78
        log_t_stdv[k] = std::log(t_stdv[k]);
        t_precision[k] = 1.0 / (t_stdv[k] * t_stdv[k]);
        log_a_stdv[k] = std::log(a_stdv[k]);
        a_precision[k] = 1.0 / (a_stdv[k] * a_stdv[k]);
      } else {
        log_t_stdv[k] = 1.0;  // sic: 1.0, not 0.0, :84
        t_stdv[k] = 1.0;
        t_precision[k] = 1.0;
        log_a_stdv[k] = 1.0;
        a_stdv[k] = 1.0;
        a_precision[k] = 1.0;
      }
    }
    t_syn.resize(ns);
    a_syn.resize(ns);
  }

  // forward_calc_travel_time_single, :142-179 (evt 0-based; xyz = the event's hypocentre)
  void travel_time_single(int32_t evt, const double xyz[3], const double* t_corr, double beta,
                          double* out) const {
    const double x = xyz[0], y = xyz[1], z = xyz[2];
    const size_t o = static_cast<size_t>(evt) * n_sta;
    for (int32_t j = 0; j < n_sta; ++j) {
      const double tc = t_corr[j];
      const double dx = x - sta_x[j], dy = y - sta_y[j], dz = z - sta_z[j];
      out[j] = std::sqrt(dx * dx + dy * dy + dz * dz) / beta - tc;
    }
    double num = 0.0, den = 0.0;  // sum() accumulates left to right
    for (int32_t j = 0; j < n_sta; ++j) num += t_precision[o + j] * (out[j] - t_obs[o + j]);
    for (int32_t j = 0; j < n_sta; ++j) den += t_precision[o + j];
    const double t_mean = num / den;
    for (int32_t j = 0; j < n_sta; ++j) out[j] = out[j] - t_mean;
  }

  // forward_calc_amp_single, :226-264
  void amp_single(int32_t evt, const double xyz[3], const double* a_corr, double q, double beta,
                  double* out) const {
    const double pi = std::acos(-1.0);
    const double freq = 5.0;
    const double x = xyz[0], y = xyz[1], z = xyz[2];
    const size_t o = static_cast<size_t>(evt) * n_sta;
    for (int32_t j = 0; j < n_sta; ++j) {
      const double ac = a_corr[j];
      const double dx = x - sta_x[j], dy = y - sta_y[j], dz = z - sta_z[j];
      const double d = std::sqrt(dx * dx + dy * dy + dz * dz);
      out[j] = -d * pi * freq / (q * beta) - std::log(d) - ac;
    }
    double num = 0.0, den = 0.0;
    for (int32_t j = 0; j < n_sta; ++j) num += a_precision[o + j] * (out[j] - a_obs[o + j]);
    for (int32_t j = 0; j < n_sta; ++j) den += a_precision[o + j];
    const double a_mean = num / den;
    for (int32_t j = 0; j < n_sta; ++j) out[j] = out[j] - a_mean;
  }

  // forward_calc_log_likelihood, :268-303.  hypo is the model vector x(3E) (x,y,z
  // interleaved per event).  The reference accumulates ALL time terms (events outer,
  // stations inner) and then ALL amplitude terms into one scalar, in that order.
  // per_event (optional) receives each event's own contribution summed separately.
  double calc_log_likelihood(const double* hypo, const double* t_corr, double vs,
                             const double* a_corr, double qs, double* per_event = nullptr) const {
    const double l2ph = log_2pi_half();
    double ll = 0.0;
    if (per_event)
      for (int32_t i = 0; i < n_events; ++i) per_event[i] = 0.0;
    if (use_time) {
      for (int32_t i = 0; i < n_events; ++i) {
        travel_time_single(i, hypo + 3 * i, t_corr, vs, t_syn.data());
        const size_t o = static_cast<size_t>(i) * n_sta;
        for (int32_t j = 0; j < n_sta; ++j) {
          const double r = t_obs[o + j] - t_syn[j];
          const double term = r * r / (2.0 * (t_stdv[o + j] * t_stdv[o + j]));
          ll = ll - term - l2ph - log_t_stdv[o + j];
          if (per_event) per_event[i] = per_event[i] - term - l2ph - log_t_stdv[o + j];
        }
      }
    }
    if (use_amp) {
      for (int32_t i = 0; i < n_events; ++i) {
        amp_single(i, hypo + 3 * i, a_corr, qs, vs, a_syn.data());
        const size_t o = static_cast<size_t>(i) * n_sta;
        for (int32_t j = 0; j < n_sta; ++j) {
          const double r = a_obs[o + j] - a_syn[j];
          const double term = r * r / (2.0 * (a_stdv[o + j] * a_stdv[o + j]));
          ll = ll - term - l2ph - log_a_stdv[o + j];
          if (per_event) per_event[i] = per_event[i] - term - l2ph - log_a_stdv[o + j];
        }
      }
    }
    return ll;
  }

  // One event's log-likelihood contribution (time block then amplitude block), the
  // quantity modes B and C cache per (chain, event).
  double event_log_likelihood(int32_t evt, const double xyz[3], const double* t_corr, double vs,
                              const double* a_corr, double qs) const {
    const double l2ph = log_2pi_half();
    const size_t o = static_cast<size_t>(evt) * n_sta;
    double ll = 0.0;
    if (use_time) {
      travel_time_single(evt, xyz, t_corr, vs, t_syn.data());
      for (int32_t j = 0; j < n_sta; ++j) {
        const double r = t_obs[o + j] - t_syn[j];
        ll = ll - r * r / (2.0 * (t_stdv[o + j] * t_stdv[o + j])) - l2ph - log_t_stdv[o + j];
      }
    }
    if (use_amp) {
      amp_single(evt, xyz, a_corr, qs, vs, a_syn.data());
      for (int32_t j = 0; j < n_sta; ++j) {
        const double r = a_obs[o + j] - a_syn[j];
        ll = ll - r * r / (2.0 * (a_stdv[o + j] * a_stdv[o + j])) - l2ph - log_a_stdv[o + j];
      }
    }
    return ll;
  }

  // forward_partially_update_log_likelihood, :307-362.  Adds the old event's station
  // terms one by one INTO the running total and subtracts the new ones (:319-357).
  double partially_update(int32_t evt, const double xyz_old[3], double ll_old,
                          const double xyz_new[3], const double* t_corr, double vs,
                          const double* a_corr, double qs) const {
    const double l2ph = log_2pi_half();
    const size_t o = static_cast<size_t>(evt) * n_sta;
    double ll = ll_old;
    if (use_time) {
      travel_time_single(evt, xyz_old, t_corr, vs, t_syn.data());
      for (int32_t j = 0; j < n_sta; ++j) {
        const double r = t_obs[o + j] - t_syn[j];
        ll = ll + r * r / (2.0 * (t_stdv[o + j] * t_stdv[o + j])) + l2ph + log_t_stdv[o + j];
      }
      travel_time_single(evt, xyz_new, t_corr, vs, t_syn.data());
      for (int32_t j = 0; j < n_sta; ++j) {
        const double r = t_obs[o + j] - t_syn[j];
        ll = ll - r * r / (2.0 * (t_stdv[o + j] * t_stdv[o + j])) - l2ph - log_t_stdv[o + j];
      }
    }
    if (use_amp) {
      amp_single(evt, xyz_old, a_corr, qs, vs, a_syn.data());
      for (int32_t j = 0; j < n_sta; ++j) {
        const double r = a_obs[o + j] - a_syn[j];
        ll = ll + r * r / (2.0 * (a_stdv[o + j] * a_stdv[o + j])) + l2ph + log_a_stdv[o + j];
      }
      amp_single(evt, xyz_new, a_corr, qs, vs, a_syn.data());
      for (int32_t j = 0; j < n_sta; ++j) {
        const double r = a_obs[o + j] - a_syn[j];
        ll = ll - r * r / (2.0 * (a_stdv[o + j] * a_stdv[o + j])) - l2ph - log_a_stdv[o + j];
      }
    }
    return ll;
  }
};

// ------------------------------------------------------------------------------------
// cls_mcmc  (src/cls_mcmc.f90) -- one joint chain.  The reference deep-copies five
// models per proposal; here the proposal is (model id, index, new value) applied in
// place and rolled back on rejection, which is arithmetically identical.
// ------------------------------------------------------------------------------------
struct Proposal {
  int32_t type = 0;    // i_proposal_type 1..7
  int32_t which = 0;   // 1 vs, 2 t_corr, 3 qs, 4 a_corr, 5 hypo
  int32_t index = 0;   // 0-based index inside that model
  int32_t evt_id = -999;  // 1-based event id for hypo moves (:138,:164)
  double x_old = 0.0, x_new = 0.0;
  double log_prior_ratio = 0.0;
  bool prior_ok = true;
};

struct Chain {
  int32_t n_events = 0, n_sta = 0;
  Model hypo, t_corr, vs, a_corr, qs;
  double temp = 1.0;
  double log_likelihood = -9.0e+300;  // init_mcmc :88
  int64_t n_propose[7] = {0, 0, 0, 0, 0, 0, 0};
  int64_t n_accept[7] = {0, 0, 0, 0, 0, 0, 0};
  int32_t i_iter = 0;
  bool is_accepted = false;
  double p_vs = 0, p_t_corr = 0, p_qs = 0, p_a_corr = 0, p_hypo = 1;

  // proposal probabilities, init_mcmc :91-108
  void set_solve(bool s_t, bool s_vs, bool s_a, bool s_qs) {
    p_vs = s_vs ? 0.025 : 0.0;
    p_t_corr = s_t ? 0.025 : 0.0;
    p_qs = s_qs ? 0.025 : 0.0;
    p_a_corr = s_a ? 0.025 : 0.0;
    p_hypo = 1.0 - p_vs - p_t_corr - p_qs - p_a_corr;
  }
  Model& model_of(int32_t which) {
    switch (which) {
      case 1: return vs;
      case 2: return t_corr;
      case 3: return qs;
      case 4: return a_corr;
      default: return hypo;
    }
  }
  // mcmc_propose_model, :115-172.  Draw order: a_select; [index]; [component]; 2 for the
  // Box-Muller step.  The proposed value is written into the model (like *_proposed).
  Proposal propose(Xorshift128& rng) {
    Proposal p;
    const double a_select = rng.rand_u();
    if (a_select < p_vs) {
      p.which = 1;
      p.index = 0;
      p.type = 1;
    } else if (a_select < p_vs + p_t_corr) {
      p.which = 2;
      p.index = static_cast<int32_t>(rng.rand_u() * n_sta);  // id-1, :145
      p.type = 2;
    } else if (a_select < p_vs + p_t_corr + p_qs) {
      p.which = 3;
      p.index = 0;
      p.type = 3;
    } else if (a_select < p_vs + p_t_corr + p_qs + p_a_corr) {
      p.which = 4;
      p.index = static_cast<int32_t>(rng.rand_u() * n_sta);  // :155
      p.type = 4;
    } else {
      const int32_t id = static_cast<int32_t>(rng.rand_u() * n_events) + 1;  // :160
      const int32_t icmp = static_cast<int32_t>(rng.rand_u() * 3);           // :161
      p.which = 5;
      p.index = 3 * id - icmp - 1;  // Fortran index 3*id-icmp, :162
      p.type = 5 + icmp;            // :163
      p.evt_id = id;
    }
    Model& m = model_of(p.which);
    p.x_old = m.x[p.index];
    const double g = rng.rand_g();
    p.x_new = m.perturb_with(p.index, g, p.log_prior_ratio, p.prior_ok);
    m.x[p.index] = p.x_new;
    return p;
  }
  // mcmc_judge_model, :176-226.  ll_new is only read when prior_ok.
  void judge(Xorshift128& rng, const Proposal& p, double ll_new) {
    const double eps = 2.220446049250313e-16;  // epsilon(1.d0)
    if (temp < 1.0 + eps) n_propose[p.type - 1] += 1;
    is_accepted = false;
    if (p.prior_ok) {
      double ratio = (ll_new - log_likelihood) / temp;
      ratio = ratio + p.log_prior_ratio;
      const double r = rng.rand_u();
      if (r >= eps) {
        if (std::log(r) <= ratio) is_accepted = true;
      }
    }
    if (is_accepted) {
      log_likelihood = ll_new;
      if (temp < 1.0 + eps) n_accept[p.type - 1] += 1;
    } else {
      model_of(p.which).x[p.index] = p.x_old;  // proposal discarded
    }
    i_iter += 1;
  }
};

}  // namespace hto
