// TEST INFRASTRUCTURE -- NOT PRODUCT CODE (see htm_oracle.hpp header).
// extern "C" surface of the oracle, loaded with ctypes by tests/, smoke() and bench.py's
// cpu_baseline / --impl reference legs only.
#include <chrono>
#include <cstdio>
#include <cstring>

#include "htm_oracle_run.hpp"
#include "htm_oracle_measure.hpp"
#include "htm_oracle_select.hpp"

using hto::Oracle;

extern "C" {

// ---- mod_random known-answer access --------------------------------------------------
void hto_rng_seeds(int32_t rank, int32_t out[4]) {
  hto::Xorshift128 g;
  g.init(5551111, 453222, 4444431, 6765, rank);
  out[0] = static_cast<int32_t>(g.x);
  out[1] = static_cast<int32_t>(g.y);
  out[2] = static_cast<int32_t>(g.z);
  out[3] = static_cast<int32_t>(g.w);
}
// kind: 0 rand_u, 1 rand_u2, 2 rand_g, 3 rand_r, 4 raw w
void hto_rng_draw(int32_t rank, int32_t kind, int32_t n, double* out) {
  hto::Xorshift128 g;
  g.init(5551111, 453222, 4444431, 6765, rank);
  for (int32_t i = 0; i < n; ++i) {
    switch (kind) {
      case 0: out[i] = g.rand_u(); break;
      case 1: out[i] = g.rand_u2(); break;
      case 2: out[i] = g.rand_g(); break;
      case 3: out[i] = g.rand_r(); break;
      default: out[i] = static_cast<double>(g.next_raw()); break;
    }
  }
}
void hto_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) {
  hto::Philox::gen(seed, c0, c1, c2, c3, out);
}

// ---- cls_model --------------------------------------------------------------------------
// model_perturb's arithmetic for one scalar: returns x_new
double hto_perturb(double x_old, double mu, double sigma, double step, int32_t prior_type, double g,
                   double* log_prior_ratio, int32_t* prior_ok) {
  hto::Model m(1);
  m.set_prior(0, mu, sigma, prior_type);
  m.step_size[0] = step;
  m.x[0] = x_old;
  bool ok;
  const double xn = m.perturb_with(0, g, *log_prior_ratio, ok);
  *prior_ok = ok ? 1 : 0;
  return xn;
}
int32_t hto_judge_swap(double t1, double t2, double l1, double l2, double r) {
  return hto::judge_swap_with(t1, t2, l1, l2, r) ? 1 : 0;
}

// ---- handle -----------------------------------------------------------------------------
void* hto_create(const htm_config* cfg, const double* sx, const double* sy, const double* sz,
                 const double* tobs, const double* tstd, const double* aobs, const double* astd,
                 const double* xmu, const double* ymu) {
  Oracle* o = new Oracle();
  o->create(*cfg, sx, sy, sz, tobs, tstd, aobs, astd, xmu, ymu);
  return o;
}
void hto_destroy(void* h) { delete static_cast<Oracle*>(h); }

void hto_set_event_offset(void* h, int32_t off) { static_cast<Oracle*>(h)->event_offset = off; }

void hto_set_rank_shard(void* h, uint32_t chain_offset, uint32_t j_total, uint32_t swap_stream) {
  Oracle* o = static_cast<Oracle*>(h);
  o->chain_offset = chain_offset;
  o->J_total = j_total;
  o->swap_stream = swap_stream;
}

// event shards of one joint ensemble (blocked Gibbs): see Oracle::sum_hook
void hto_set_sum_hook(void* h, void (*hook)(double*, void*), void* user) {
  Oracle* o = static_cast<Oracle*>(h);
  o->sum_hook = hook;
  o->sum_hook_user = user;
}

void hto_set_globals(void* h, double vs, double qs, const double* tc, const double* ac) {
  Oracle* o = static_cast<Oracle*>(h);
  o->fixed_vs = vs;
  o->fixed_qs = qs;
  o->fixed_t_corr.assign(tc, tc + o->S);
  o->fixed_a_corr.assign(ac, ac + o->S);
}

// forward%calc_log_likelihood for M models; per_event may be NULL ([M][E])
void hto_loglik(void* h, int32_t M, const double* hypo, const double* tc, const double* ac,
                const double* vs, const double* qs, double* L, double* per_event) {
  Oracle* o = static_cast<Oracle*>(h);
  for (int32_t m = 0; m < M; ++m) {
    L[m] = o->fwd.calc_log_likelihood(hypo + static_cast<size_t>(m) * 3 * o->E,
                                      tc + static_cast<size_t>(m) * o->S, vs[m],
                                      ac + static_cast<size_t>(m) * o->S, qs[m],
                                      per_event ? per_event + static_cast<size_t>(m) * o->E : nullptr);
  }
}
// forward%partially_update_log_likelihood (evt 1-based as in the reference)
double hto_partial_update(void* h, int32_t evt_id, const double* hypo_old, double ll_old,
                          const double* hypo_new, const double* tc, double vs, const double* ac,
                          double qs) {
  Oracle* o = static_cast<Oracle*>(h);
  const int32_t e = evt_id - 1;
  return o->fwd.partially_update(e, hypo_old + 3 * e, ll_old, hypo_new + 3 * e, tc, vs, ac, qs);
}
// the precomputed tables of init_forward (for the degenerate-sigma checks)
void hto_forward_tables(void* h, double* t_prec, double* log_t, double* a_prec, double* log_a) {
  Oracle* o = static_cast<Oracle*>(h);
  const size_t n = static_cast<size_t>(o->S) * o->E;
  std::memcpy(t_prec, o->fwd.t_precision.data(), n * sizeof(double));
  std::memcpy(log_t, o->fwd.log_t_stdv.data(), n * sizeof(double));
  std::memcpy(a_prec, o->fwd.a_precision.data(), n * sizeof(double));
  std::memcpy(log_a, o->fwd.log_a_stdv.data(), n * sizeof(double));
}

// ---- chains -----------------------------------------------------------------------------
void hto_record_draws(void* h, int32_t on) { static_cast<Oracle*>(h)->record_draws(on != 0); }

void hto_init_chains(void* h) {
  Oracle* o = static_cast<Oracle*>(h);
  if (o->cfg.mode == HTM_MODE_FACTORISED)
    o->init_chains_factorised();
  else if (o->cfg.mode == HTM_MODE_BLOCKED_GIBBS)
    o->init_chains_gibbs();
  else
    o->init_chains_reference();
}

void hto_get_chain_state(void* h, int32_t r, int32_t j, double* hypo, double* tc, double* ac,
                         double* vs, double* qs, double* temp, double* L) {
  Oracle* o = static_cast<Oracle*>(h);
  if (o->cfg.mode == HTM_MODE_FACTORISED) {
    double lsum = 0.0;
    for (int32_t e = 0; e < o->E; ++e) {
      const size_t i = o->bidx(e, r, j);
      hypo[3 * e] = o->bx[i];
      hypo[3 * e + 1] = o->by[i];
      hypo[3 * e + 2] = o->bz[i];
      lsum += o->bL[i];
    }
    if (tc) std::memcpy(tc, o->fixed_t_corr.data(), o->S * sizeof(double));
    if (ac) std::memcpy(ac, o->fixed_a_corr.data(), o->S * sizeof(double));
    *vs = o->fixed_vs;
    *qs = o->fixed_qs;
    *temp = o->bT[o->bidx(0, r, j)];
    *L = lsum;
    return;
  }
  if (o->cfg.mode == HTM_MODE_BLOCKED_GIBBS) {
    const Oracle::GibbsChain& g = o->gc[static_cast<size_t>(r) * o->n_chains + j];
    for (int32_t e = 0; e < o->E; ++e) {
      hypo[3 * e] = g.x[e];
      hypo[3 * e + 1] = g.y[e];
      hypo[3 * e + 2] = g.z[e];
    }
    if (tc) std::memcpy(tc, g.tc.data(), o->S * sizeof(double));
    if (ac) std::memcpy(ac, g.ac.data(), o->S * sizeof(double));
    *vs = g.vs;
    *qs = g.qs;
    *temp = g.temp;
    *L = g.L;
    return;
  }
  const hto::Chain& c = o->chain(r, j);
  std::memcpy(hypo, c.hypo.x.data(), c.hypo.x.size() * sizeof(double));
  if (tc) std::memcpy(tc, c.t_corr.x.data(), o->S * sizeof(double));
  if (ac) std::memcpy(ac, c.a_corr.x.data(), o->S * sizeof(double));
  *vs = c.vs.x[0];
  *qs = c.qs.x[0];
  *temp = c.temp;
  *L = c.log_likelihood;
}
// mode B per-chain arrays [event][rank][chain]
void hto_get_factorised_state(void* h, double* x, double* y, double* z, double* L, double* T) {
  Oracle* o = static_cast<Oracle*>(h);
  const size_t n = o->bx.size();
  std::memcpy(x, o->bx.data(), n * sizeof(double));
  std::memcpy(y, o->by.data(), n * sizeof(double));
  std::memcpy(z, o->bz.data(), n * sizeof(double));
  std::memcpy(L, o->bL.data(), n * sizeof(double));
  std::memcpy(T, o->bT.data(), n * sizeof(double));
}

// ---- runs ---------------------------------------------------------------------------------
void hto_run(void* h, int32_t iter_first, int32_t iter_last, htm_step_trace* trace,
             htm_swap_trace* swaps) {
  Oracle* o = static_cast<Oracle*>(h);
  if (o->cfg.mode == HTM_MODE_FACTORISED)
    o->run_factorised(iter_first, iter_last, trace, swaps);
  else if (o->cfg.mode == HTM_MODE_BLOCKED_GIBBS)
    o->run_gibbs(iter_first, iter_last, trace, swaps);
  else
    o->run_reference(iter_first, iter_last, trace, swaps);
}
// mode A, one thread per virtual rank; returns wall seconds
double hto_run_threaded(void* h, int32_t iter_first, int32_t iter_last) {
  Oracle* o = static_cast<Oracle*>(h);
  const auto t0 = std::chrono::steady_clock::now();
  o->run_reference_threaded(iter_first, iter_last);
  const auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

int64_t hto_n_draws(void* h, int32_t rank) {
  return static_cast<int64_t>(static_cast<Oracle*>(h)->tapes[rank].size());
}
void hto_get_draws(void* h, int32_t rank, int32_t* buf) {
  Oracle* o = static_cast<Oracle*>(h);
  std::memcpy(buf, o->tapes[rank].data(), o->tapes[rank].size() * sizeof(int32_t));
}
void hto_clear_draws(void* h) {
  Oracle* o = static_cast<Oracle*>(h);
  for (auto& t : o->tapes) t.clear();
}

void hto_get_counts(void* h, int64_t np[7], int64_t na[7]) { static_cast<Oracle*>(h)->counts(np, na); }

int32_t hto_n_samples(void* h, int32_t rank) {
  Oracle* o = static_cast<Oracle*>(h);
  return static_cast<int32_t>(o->out[rank].iter.size() - o->out[rank].fetched);
}
// same contract as htm_fetch_samples
void hto_fetch_samples(void* h, int32_t rank, int32_t max_records, int32_t* n_records, int32_t* iter,
                       double* vs, double* qs, double* hypo, double* tc, double* ac) {
  Oracle* o = static_cast<Oracle*>(h);
  hto::RankOutput& ro = o->out[rank];
  size_t n = ro.iter.size() - ro.fetched;
  if (n > static_cast<size_t>(max_records)) n = max_records;
  const size_t f = ro.fetched, E3 = 3 * static_cast<size_t>(o->E), S = o->S;
  if (iter) std::memcpy(iter, ro.iter.data() + f, n * sizeof(int32_t));
  if (vs) std::memcpy(vs, ro.vs.data() + f, n * sizeof(double));
  if (qs) std::memcpy(qs, ro.qs.data() + f, n * sizeof(double));
  if (hypo) std::memcpy(hypo, ro.hypo.data() + f * E3, n * E3 * sizeof(double));
  if (tc) std::memcpy(tc, ro.t_corr.data() + f * S, n * S * sizeof(double));
  if (ac) std::memcpy(ac, ro.a_corr.data() + f * S, n * S * sizeof(double));
  ro.fetched += n;
  *n_records = static_cast<int32_t>(n);
}
int32_t hto_n_likelihood(void* h, int32_t rank) {
  Oracle* o = static_cast<Oracle*>(h);
  return static_cast<int32_t>(o->out[rank].lik_iter.size() - o->out[rank].lik_fetched);
}
void hto_fetch_likelihood(void* h, int32_t rank, int32_t max_records, int32_t* n_records,
                          int32_t* iter, double* lik) {
  Oracle* o = static_cast<Oracle*>(h);
  hto::RankOutput& ro = o->out[rank];
  size_t n = ro.lik_iter.size() - ro.lik_fetched;
  if (n > static_cast<size_t>(max_records)) n = max_records;
  if (iter) std::memcpy(iter, ro.lik_iter.data() + ro.lik_fetched, n * sizeof(int32_t));
  if (lik) std::memcpy(lik, ro.lik.data() + ro.lik_fetched, n * sizeof(double));
  ro.lik_fetched += n;
  *n_records = static_cast<int32_t>(n);
}

// hypo_tremor_select for n_events windows: out [n_events][6] = vs, t0, b, a0, cc_t, cc_a; selected [n_events]
void hto_select(int32_t S, int32_t E, const double* sta_x, const double* sta_y, const double* sta_z, double z_guess,
                const double* t, const double* t_err, const double* a, const double* a_err, double vs_min, double vs_max,
                double b_min, double b_max, double* out, int32_t* selected) {
  for (int32_t e = 0; e < E; ++e) {
    const size_t o = static_cast<size_t>(e) * S;
    double* r = out + static_cast<size_t>(e) * 6;
    hto::select_window(S, sta_x, sta_y, sta_z, z_guess, t + o, t_err + o, a + o, a_err + o, r);
    selected[e] = (r[0] >= vs_min && r[0] <= vs_max && r[2] >= b_min && r[2] <= b_max) ? 1 : 0;
  }
}

// hypo_tremor_measure's lag / amplitude optimisation for n_win windows cut from the merged envelopes env [S][n_total]
// (src/cls_measurer.f90:331-343: window id -> samples (id - 1) * n_step + 1 ... + n); t, t_stdv, amp, amp_stdv
// [n_win][S]; lag (may be null) [n_win][S (S - 1) / 2]
void hto_measure(int32_t S, int64_t n_total, const double* env, double dt, int32_t n, int32_t n_step, int32_t n_win,
                 const int32_t* win_id, double* t, double* t_stdv, double* amp, double* amp_stdv, int32_t* lag) {
  std::vector<double> x(static_cast<size_t>(S) * n);
  const size_t P = static_cast<size_t>(S) * (S - 1) / 2;
  for (int32_t w = 0; w < n_win; ++w) {
    const int64_t j1 = static_cast<int64_t>(win_id[w] - 1) * n_step;
    for (int32_t i = 0; i < S; ++i)
      for (int32_t m = 0; m < n; ++m) x[static_cast<size_t>(i) * n + m] = env[static_cast<size_t>(i) * n_total + j1 + m];
    const size_t o = static_cast<size_t>(w) * S;
    hto::measure_window(S, n, dt, x.data(), t + o, t_stdv + o, amp + o, amp_stdv + o, lag ? lag + w * P : nullptr);
  }
}

void hto_detect(int32_t S, int64_t n_total, const double* env, int32_t n, int32_t n_step, double alpha, int32_t n_pair_thred,
                int32_t n_win, double* cc_thred, double* cc_max, int32_t* detected, int32_t* n_above) {
  hto::detect_windows(S, static_cast<long>(n_total), env, n, n_step, alpha, n_pair_thred, n_win, cc_thred, cc_max, detected,
                      n_above);
}

}  // extern "C"
