// TEST INFRASTRUCTURE -- NOT PRODUCT CODE (see htm_oracle.hpp header).
//
// The driver half of the oracle: chain set-up, the main loop, recording, the tempering
// swap (src/hypo_tremor_mcmc.f90:72-291, src/cls_parallel.f90:100-302) with MPI ranks
// turned into *virtual ranks* (one xorshift128 stream each, chains of a rank stepped
// sequentially, one swap per iteration), plus CPU statements of the two B200 schedules
// (factorised, blocked Gibbs) on Philox draws so the GPU kernels can be checked step by
// step against reference-restated arithmetic.
#pragma once
#include <atomic>
#include <cstring>
#include <thread>

#include "htm_oracle.hpp"

namespace hto {

static const double kEps = 2.220446049250313e-16;  // epsilon(1.d0)

// What one rank appended to its six .out files (src/hypo_tremor_mcmc.f90:270-280).
struct RankOutput {
  std::vector<int32_t> iter;      // one per sample record
  std::vector<double> vs, qs;     // one per record
  std::vector<double> hypo;       // 3E per record
  std::vector<double> t_corr;     // S per record
  std::vector<double> a_corr;     // S per record
  std::vector<int32_t> lik_iter;  // likelihood file (includes burn-in)
  std::vector<double> lik;
  size_t fetched = 0, lik_fetched = 0;
};

// judge_swap, src/cls_parallel.f90:285-302, for a given uniform r
static inline bool judge_swap_with(double temp1, double temp2, double l1, double l2, double r) {
  const double del_s = (l2 - l1) * (1.0 / temp1 - 1.0 / temp2);
  bool acc = false;
  if (r >= kEps) {
    if (std::log(r) <= del_s) acc = true;
  }
  return acc;
}

// sense-reversing spin barrier for the threaded mode-A run (stands in for the blocking
// MPI_Bcast / Send / Recv every iteration, src/cls_parallel.f90:112,154-203)
struct SpinBarrier {
  std::atomic<int> count{0};
  std::atomic<int> sense{0};
  int n = 1;
  void wait(int& local_sense) {
    local_sense ^= 1;
    if (count.fetch_add(1, std::memory_order_acq_rel) == n - 1) {
      count.store(0, std::memory_order_relaxed);
      sense.store(local_sense, std::memory_order_release);
    } else {
      while (sense.load(std::memory_order_acquire) != local_sense) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
      }
    }
  }
};

struct Oracle {
  htm_config cfg;
  Forward fwd;
  std::vector<double> x_mu, y_mu;
  int32_t n_ranks = 1, n_chains = 1, E = 0, S = 0;
  std::vector<Xorshift128> rng;             // one stream per virtual rank
  std::vector<std::vector<int32_t>> tapes;  // recorded raw draws per rank
  std::vector<Chain> chains;                // [rank][chain], mode A and C
  std::vector<RankOutput> out;              // per rank
  // fixed globals for mode B
  double fixed_vs = 0, fixed_qs = 0;
  std::vector<double> fixed_t_corr, fixed_a_corr;
  // mode B state: [event][rank][chain]
  std::vector<double> bx, by, bz, bL, bT;
  int64_t b_propose[7] = {0, 0, 0, 0, 0, 0, 0}, b_accept[7] = {0, 0, 0, 0, 0, 0, 0};
  int32_t event_offset = 0;  // global id of this shard's first event

  Chain& chain(int32_t r, int32_t j) { return chains[static_cast<size_t>(r) * n_chains + j]; }

  void create(const htm_config& c, const double* sx, const double* sy, const double* sz,
              const double* tobs, const double* tstd, const double* aobs, const double* astd,
              const double* xm, const double* ym) {
    cfg = c;
    n_ranks = c.n_procs;
    n_chains = c.n_chains;
    E = c.n_events;
    S = c.n_sta;
    fwd.init(S, E, sx, sy, sz, tobs, tstd, aobs, astd, c.use_amp != 0, c.use_time != 0);
    x_mu.assign(xm, xm + E);
    y_mu.assign(ym, ym + E);
    rng.assign(n_ranks, Xorshift128());
    tapes.assign(n_ranks, {});
    out.assign(n_ranks, RankOutput());
    fixed_vs = c.prior_vs;
    fixed_qs = c.prior_qs;
    fixed_t_corr.assign(S, c.prior_t_corr);
    fixed_a_corr.assign(S, c.prior_a_corr);
  }
  void record_draws(bool on) {
    for (int32_t r = 0; r < n_ranks; ++r) rng[r].tape = on ? &tapes[r] : nullptr;
  }

  // ---- chain set-up, src/hypo_tremor_mcmc.f90:72,120-211 ---------------------------------
  Chain make_chain_models() const {
    Chain c;
    c.n_events = E;
    c.n_sta = S;
    c.t_corr = Model(S);
    c.a_corr = Model(S);
    c.hypo = Model(3 * E);
    c.vs = Model(1);
    c.qs = Model(1);
    for (int32_t i = 0; i < S; ++i) {
      c.t_corr.set_prior(i, cfg.prior_t_corr, cfg.prior_width_t_corr);
      c.t_corr.step_size[i] = cfg.step_size_t_corr;
      c.t_corr.x[i] = cfg.prior_t_corr;
      c.a_corr.set_prior(i, cfg.prior_a_corr, cfg.prior_width_a_corr);
      c.a_corr.step_size[i] = cfg.step_size_a_corr;
      c.a_corr.x[i] = cfg.prior_a_corr;
    }
    for (int32_t i = 0; i < E; ++i) {
      c.hypo.set_prior(3 * i, x_mu[i], cfg.prior_width_xy);
      c.hypo.set_prior(3 * i + 1, y_mu[i], cfg.prior_width_xy);
      c.hypo.set_prior(3 * i + 2, cfg.prior_z, cfg.prior_width_z, 1);
      c.hypo.step_size[3 * i] = cfg.step_size_xy;
      c.hypo.step_size[3 * i + 1] = cfg.step_size_xy;
      c.hypo.step_size[3 * i + 2] = cfg.step_size_z;
    }
    c.vs.set_prior(0, cfg.prior_vs, cfg.prior_width_vs);
    c.vs.step_size[0] = cfg.step_size_vs;
    c.vs.x[0] = cfg.prior_vs;
    c.qs.set_prior(0, cfg.prior_qs, cfg.prior_width_qs);
    c.qs.step_size[0] = cfg.step_size_qs;
    c.qs.x[0] = cfg.prior_qs;
    c.set_solve(cfg.solve_t_corr != 0, cfg.solve_vs != 0, cfg.solve_a_corr != 0, cfg.solve_qs != 0);
    return c;
  }

  void init_chains_reference() {
    chains.assign(static_cast<size_t>(n_ranks) * n_chains, Chain());
    for (int32_t r = 0; r < n_ranks; ++r) {
      rng[r].init(5551111, 453222, 4444431, 6765, r);  // :72
      for (int32_t j = 0; j < n_chains; ++j) {
        Chain c = make_chain_models();
        if (cfg.solve_t_corr) c.t_corr.generate(rng[r]);  // :126-132
        if (cfg.solve_a_corr) c.a_corr.generate(rng[r]);  // :143-149
        c.hypo.generate(rng[r]);                          // :171
        if (j + 1 <= cfg.n_cool) {                        // :202-208
          c.temp = 1.0;
        } else {
          c.temp = std::exp((rng[r].rand_u() * (1.0 - kEps) + kEps) * std::log(cfg.temp_high));
        }
        chain(r, j) = std::move(c);
      }
    }
  }

  // ---- loop body for one (iteration, rank, chain), src/hypo_tremor_mcmc.f90:238-281 -------
  void step_chain(int32_t r, int32_t j, int32_t it, htm_step_trace* tr) {
    step_chain_with(fwd, r, j, it, tr);
  }
  // f: the Forward whose scratch arrays this caller owns (per-thread copy when threaded)
  void step_chain_with(const Forward& f, int32_t r, int32_t j, int32_t it, htm_step_trace* tr) {
    Chain& mc = chain(r, j);
    Xorshift128& g = rng[r];
    Proposal p = mc.propose(g);
    double ll_new = 0.0;
    if (p.prior_ok) {
      if (p.evt_id > 0 && it > 1) {  // :246
        const int32_t e = p.evt_id - 1;
        double xyz_new[3] = {mc.hypo.x[3 * e], mc.hypo.x[3 * e + 1], mc.hypo.x[3 * e + 2]};
        double xyz_old[3] = {xyz_new[0], xyz_new[1], xyz_new[2]};
        xyz_old[p.index - 3 * e] = p.x_old;
        ll_new = f.partially_update(e, xyz_old, mc.log_likelihood, xyz_new, mc.t_corr.x.data(),
                                    mc.vs.x[0], mc.a_corr.x.data(), mc.qs.x[0]);
      } else {
        ll_new = f.calc_log_likelihood(mc.hypo.x.data(), mc.t_corr.x.data(), mc.vs.x[0],
                                       mc.a_corr.x.data(), mc.qs.x[0]);
      }
    }
    mc.judge(g, p, ll_new);
    if (tr) {
      tr->proposal_type = p.type;
      tr->index = p.index + 1;
      tr->prior_ok = p.prior_ok ? 1 : 0;
      tr->accepted = mc.is_accepted ? 1 : 0;
      tr->log_likelihood = mc.log_likelihood;
    }
    // Recording, :270-280.  mod(i, n_interval) == 1 is never true for n_interval = 1.
    if (mc.temp < 1.0 + kEps && (it % cfg.n_interval) == 1) {
      RankOutput& o = out[r];
      if (it > cfg.n_burn) {
        o.iter.push_back(it);
        o.vs.push_back(mc.vs.x[0]);
        o.hypo.insert(o.hypo.end(), mc.hypo.x.begin(), mc.hypo.x.end());
        o.t_corr.insert(o.t_corr.end(), mc.t_corr.x.begin(), mc.t_corr.x.end());
        o.qs.push_back(mc.qs.x[0]);
        o.a_corr.insert(o.a_corr.end(), mc.a_corr.x.begin(), mc.a_corr.x.end());
      }
      o.lik_iter.push_back(it);
      o.lik.push_back(mc.log_likelihood);
    }
  }

  // ---- parallel_swap_temperature + select_pair, src/cls_parallel.f90:100-240 ---------------
  void swap_temperature(htm_swap_trace* tr) {
    Xorshift128& g0 = rng[0];
    // select_pair on rank 0's stream, :226-234.  rand_u()*n_proc*n_chain evaluates
    // left to right in float64.
    const int32_t i1 = static_cast<int32_t>(g0.rand_u() * n_ranks * n_chains);
    int32_t i2;
    for (;;) {
      i2 = static_cast<int32_t>(g0.rand_u() * n_ranks * n_chains);
      if (i1 != i2) break;
    }
    const int32_t rank1 = i1 / n_chains, rank2 = i2 / n_chains;
    const int32_t chain1 = i1 % n_chains, chain2 = i2 % n_chains;  // 0-based here
    Chain& mc1 = chain(rank1, chain1);
    Chain& mc2 = chain(rank2, chain2);
    // judge_swap draws on rank1's stream whether or not rank1 == rank2, :129,:163
    const double r = rng[rank1].rand_u();
    const bool acc = judge_swap_with(mc1.temp, mc2.temp, mc1.log_likelihood, mc2.log_likelihood, r);
    if (acc) {
      const double t1 = mc1.temp;
      mc1.temp = mc2.temp;
      mc2.temp = t1;
    }
    if (tr) {
      tr->rank1 = rank1;
      tr->chain1 = chain1 + 1;
      tr->rank2 = rank2;
      tr->chain2 = chain2 + 1;
      tr->accepted = acc ? 1 : 0;
      tr->reserved = 0;
    }
  }

  // ---- mode A main loop, src/hypo_tremor_mcmc.f90:236-284 ----------------------------------
  void run_reference(int32_t iter_first, int32_t iter_last, htm_step_trace* trace,
                     htm_swap_trace* swaps) {
    size_t k = 0;
    for (int32_t it = iter_first; it <= iter_last; ++it) {
      for (int32_t r = 0; r < n_ranks; ++r)
        for (int32_t j = 0; j < n_chains; ++j, ++k) step_chain(r, j, it, trace ? trace + k : nullptr);
      if (n_ranks * n_chains >= 2) swap_temperature(swaps ? swaps + (it - iter_first) : nullptr);
    }
  }

  // Same loop, one std::thread per virtual rank meeting at the swap -- the CPU baseline's
  // stand-in for `mpirun -np n_procs`.  Results are identical to run_reference because a
  // rank only touches its own stream and chains between the barriers.
  void run_reference_threaded(int32_t iter_first, int32_t iter_last) {
    if (n_ranks == 1) {
      run_reference(iter_first, iter_last, nullptr, nullptr);
      return;
    }
    SpinBarrier bar;
    bar.n = n_ranks;
    // each thread needs private scratch in Forward (t_syn/a_syn are mutable members)
    std::vector<Forward> fw(n_ranks, fwd);
    auto worker = [&](int32_t r) {
      int sense = 0;
      Forward& f = fw[r];
      for (int32_t it = iter_first; it <= iter_last; ++it) {
        for (int32_t j = 0; j < n_chains; ++j) step_chain_with(f, r, j, it, nullptr);
        bar.wait(sense);
        if (r == 0 && n_ranks * n_chains >= 2) swap_temperature(nullptr);
        bar.wait(sense);
      }
    };
    std::vector<std::thread> th;
    for (int32_t r = 1; r < n_ranks; ++r) th.emplace_back(worker, r);
    worker(0);
    for (auto& t : th) t.join();
  }

  // parallel_output_proposal's reduction, src/cls_parallel.f90:259-268
  void counts(int64_t np[7], int64_t na[7]) const {
    for (int k = 0; k < 7; ++k) np[k] = na[k] = 0;
    if (cfg.mode == HTM_MODE_FACTORISED || cfg.mode == HTM_MODE_BLOCKED_GIBBS) {
      for (int k = 0; k < 7; ++k) {
        np[k] = b_propose[k];
        na[k] = b_accept[k];
      }
      return;
    }
    for (const Chain& c : chains)
      for (int k = 0; k < 7; ++k) {
        np[k] += c.n_propose[k];
        na[k] += c.n_accept[k];
      }
  }

  // =========================================================================================
  // Mode B (factorised) on Philox draws: the schedule of the B200 kernels, built from the
  // reference-restated pieces (Forward::event_log_likelihood, Model::perturb_with's rule,
  // Chain::judge's rule, judge_swap_with).
  // =========================================================================================
  inline size_t bidx(int32_t e, int32_t r, int32_t k) const {
    return (static_cast<size_t>(e) * n_ranks + r) * n_chains + k;
  }
  inline uint32_t gid(int32_t e, int32_t r, int32_t k) const {
    return (static_cast<uint32_t>(e + event_offset) * n_ranks + r) * n_chains + k;
  }
  static inline double gauss(uint32_t wa, uint32_t wb) {
    const double pi2 = 2.0 * std::acos(-1.0);
    return std::sqrt(-2.0 * std::log(Philox::u_oo(wa))) * std::cos(pi2 * Philox::u_oo(wb));
  }
  double hot_temperature(int32_t k, uint32_t g) const {
    if (k < cfg.n_cool) return 1.0;
    if (cfg.ladder == HTM_LADDER_GEOMETRIC) {
      const int32_t n_hot = n_chains - cfg.n_cool;
      return std::exp(std::log(cfg.temp_high) * static_cast<double>(k - cfg.n_cool + 1) /
                      static_cast<double>(n_hot));
    }
    uint32_t w[4];
    Philox::gen(cfg.seed, 0u, g, PHX_TEMP, 0u, w);
    return std::exp((Philox::u_co(w[0]) * (1.0 - kEps) + kEps) * std::log(cfg.temp_high));
  }
  void init_chains_factorised() {
    const size_t n = static_cast<size_t>(E) * n_ranks * n_chains;
    bx.assign(n, 0);
    by.assign(n, 0);
    bz.assign(n, 0);
    bL.assign(n, 0);
    bT.assign(n, 1);
    for (int32_t e = 0; e < E; ++e)
      for (int32_t r = 0; r < n_ranks; ++r)
        for (int32_t k = 0; k < n_chains; ++k) {
          const size_t i = bidx(e, r, k);
          const uint32_t g = gid(e, r, k);
          uint32_t a[4], b[4];
          Philox::gen(cfg.seed, 0u, g, PHX_INIT, 0u, a);
          Philox::gen(cfg.seed, 1u, g, PHX_INIT, 0u, b);
          bx[i] = x_mu[e] + gauss(a[0], a[1]) * cfg.prior_width_xy;
          by[i] = y_mu[e] + gauss(a[2], a[3]) * cfg.prior_width_xy;
          bz[i] = cfg.prior_z + std::sqrt(-2.0 * std::log(Philox::u_oo(b[0]))) * cfg.prior_width_z;
          bT[i] = hot_temperature(k, g);
          const double xyz[3] = {bx[i], by[i], bz[i]};
          bL[i] = fwd.event_log_likelihood(e, xyz, fixed_t_corr.data(), fixed_vs,
                                           fixed_a_corr.data(), fixed_qs);
        }
  }
  // one Metropolis step of chain (e,r,k) at iteration it
  void step_factorised(int32_t e, int32_t r, int32_t k, int32_t it, htm_step_trace* tr) {
    const size_t i = bidx(e, r, k);
    uint32_t w[4];
    Philox::gen(cfg.seed, static_cast<uint32_t>(it), gid(e, r, k), PHX_STEP, 0u, w);
    const int32_t icmp = static_cast<int32_t>(Philox::below(w[0], 3));  // 0->z,1->y,2->x
    const int32_t comp = 2 - icmp;
    const double g = gauss(w[1], w[2]);
    double xyz[3] = {bx[i], by[i], bz[i]};
    const double x_old = xyz[comp];
    const double mu = comp == 0 ? x_mu[e] : (comp == 1 ? y_mu[e] : cfg.prior_z);
    const double sigma = comp == 2 ? cfg.prior_width_z : cfg.prior_width_xy;
    const double step = comp == 2 ? cfg.step_size_z : cfg.step_size_xy;
    // model_perturb's rule, src/cls_model.f90:170-187
    const double x_new = x_old + g * step;
    double lpr = -((x_new - mu) * (x_new - mu) - (x_old - mu) * (x_old - mu)) / (2.0 * sigma * sigma);
    bool prior_ok = true;
    if (comp == 2) {
      if (x_new <= mu) {
        prior_ok = false;
      } else {
        lpr = lpr + std::log(x_new - mu) - std::log(x_old - mu);
      }
    }
    // mcmc_judge_model's rule, src/cls_mcmc.f90:186-219
    const int32_t type = 5 + icmp;
    const bool cold = bT[i] < 1.0 + kEps;
    if (cold) b_propose[type - 1] += 1;
    bool acc = false;
    double ll_new = bL[i];
    if (prior_ok) {
      xyz[comp] = x_new;
      ll_new = fwd.event_log_likelihood(e, xyz, fixed_t_corr.data(), fixed_vs, fixed_a_corr.data(),
                                        fixed_qs);
      const double ratio = (ll_new - bL[i]) / bT[i] + lpr;
      const double rr = Philox::u_co(w[3]);
      if (rr >= kEps && std::log(rr) <= ratio) acc = true;
    }
    if (acc) {
      bx[i] = xyz[0];
      by[i] = xyz[1];
      bz[i] = xyz[2];
      bL[i] = ll_new;
      if (cold) b_accept[type - 1] += 1;
    }
    if (tr) {
      tr->proposal_type = type;
      tr->index = 3 * (e + 1) - icmp;
      tr->prior_ok = prior_ok ? 1 : 0;
      tr->accepted = acc ? 1 : 0;
      tr->log_likelihood = bL[i];
    }
  }
  // one swap attempt inside the tempering group (e,r)
  void swap_factorised(int32_t e, int32_t r, int32_t it, htm_swap_trace* tr) {
    if (n_chains < 2) return;
    uint32_t w[4];
    const uint32_t group = static_cast<uint32_t>(e + event_offset) * n_ranks + r;
    Philox::gen(cfg.seed, static_cast<uint32_t>(it), group, PHX_SWAP, 0u, w);
    const int32_t i1 = static_cast<int32_t>(Philox::below(w[0], n_chains));
    const int32_t i2 = (i1 + 1 + static_cast<int32_t>(Philox::below(w[1], n_chains - 1))) % n_chains;
    const size_t a = bidx(e, r, i1), b = bidx(e, r, i2);
    const bool acc = judge_swap_with(bT[a], bT[b], bL[a], bL[b], Philox::u_co(w[2]));
    if (acc) std::swap(bT[a], bT[b]);
    if (tr) {
      tr->rank1 = r;
      tr->chain1 = i1 + 1;
      tr->rank2 = r;
      tr->chain2 = i2 + 1;
      tr->accepted = acc ? 1 : 0;
      tr->reserved = 0;
    }
  }
  // Recording in factorised mode: per rank and recorded iteration, n_cool records; record m
  // holds, for every event, the m-th cold chain (in chain order) of that event's group.
  void record_factorised(int32_t it) {
    if ((it % cfg.n_interval) != 1) return;
    for (int32_t r = 0; r < n_ranks; ++r) {
      RankOutput& o = out[r];
      for (int32_t m = 0; m < cfg.n_cool; ++m) {
        std::vector<double> h(3 * static_cast<size_t>(E));
        double lsum = 0.0;
        for (int32_t e = 0; e < E; ++e) {
          int32_t seen = 0, kk = -1;
          for (int32_t k = 0; k < n_chains; ++k)
            if (bT[bidx(e, r, k)] < 1.0 + kEps) {
              if (seen == m) {
                kk = k;
                break;
              }
              ++seen;
            }
          const size_t i = bidx(e, r, kk);
          h[3 * e] = bx[i];
          h[3 * e + 1] = by[i];
          h[3 * e + 2] = bz[i];
          lsum += bL[i];
        }
        if (it > cfg.n_burn) {
          o.iter.push_back(it);
          o.vs.push_back(fixed_vs);
          o.qs.push_back(fixed_qs);
          o.hypo.insert(o.hypo.end(), h.begin(), h.end());
          o.t_corr.insert(o.t_corr.end(), fixed_t_corr.begin(), fixed_t_corr.end());
          o.a_corr.insert(o.a_corr.end(), fixed_a_corr.begin(), fixed_a_corr.end());
        }
        o.lik_iter.push_back(it);
        o.lik.push_back(lsum);
      }
    }
  }
  // trace layout: [iteration][event][rank][chain]; swaps: [iteration][event][rank]
  void run_factorised(int32_t iter_first, int32_t iter_last, htm_step_trace* trace,
                      htm_swap_trace* swaps) {
    const size_t per_it = static_cast<size_t>(E) * n_ranks * n_chains;
    const size_t g_per_it = static_cast<size_t>(E) * n_ranks;
    for (int32_t it = iter_first; it <= iter_last; ++it) {
      const size_t o = static_cast<size_t>(it - iter_first);
      for (int32_t e = 0; e < E; ++e)
        for (int32_t r = 0; r < n_ranks; ++r)
          for (int32_t k = 0; k < n_chains; ++k)
            step_factorised(e, r, k, it, trace ? trace + o * per_it + bidx(e, r, k) : nullptr);
      record_factorised(it);
      for (int32_t e = 0; e < E; ++e)
        for (int32_t r = 0; r < n_ranks; ++r)
          swap_factorised(e, r, it, swaps ? swaps + o * g_per_it + static_cast<size_t>(e) * n_ranks + r : nullptr);
    }
  }

  // =========================================================================================
  // Mode C (blocked Gibbs) on Philox draws: the schedule of the B200 joint-chain kernels.
  // J = n_ranks*n_chains joint chains (c = r*n_chains + k), each with its own vs, qs, t_corr,
  // a_corr, temperature and all E hypocentres.  One iteration of a chain:
  //   1. every event proposes one hypocentre coordinate (mode B's step, judged with the CHAIN's
  //      temperature on the event's own log-likelihood -- valid because the joint density
  //      factorises over events given the shared parameters);
  //   2. one shared parameter (uniformly among the solved ones; a random station for t_corr /
  //      a_corr) is proposed and judged on the sum over all events (cls_mcmc.f90:193-203);
  //   3. record (T = 1, mod(it, n_interval) == 1), 4. one swap attempt over all J chains
  //      (cls_parallel.f90:220-240, 285-302).
  // =========================================================================================
  struct GibbsChain {
    double vs, qs, temp, L;
    std::vector<double> tc, ac, x, y, z, Le;
  };
  std::vector<GibbsChain> gc;
  // shards of virtual ranks: Philox ids stay global so every shard draws its own streams
  uint32_t chain_offset = 0, J_total = 0, swap_stream = 0;
  int32_t J() const { return n_ranks * n_chains; }
  uint32_t Jt() const { return J_total ? J_total : static_cast<uint32_t>(J()); }
  inline uint32_t gidC(int32_t e, int32_t c) const {
    return static_cast<uint32_t>(e + event_offset) * Jt() + chain_offset + static_cast<uint32_t>(c);
  }
  void solved_types(int32_t out[4], int32_t& n) const {
    n = 0;
    if (cfg.solve_vs) out[n++] = 1;      // order of cls_mcmc.f90:139-157
    if (cfg.solve_t_corr) out[n++] = 2;
    if (cfg.solve_qs) out[n++] = 3;
    if (cfg.solve_a_corr) out[n++] = 4;
  }
  void init_chains_gibbs() {
    gc.assign(J(), GibbsChain());
    for (int32_t c = 0; c < J(); ++c) {
      GibbsChain& g = gc[c];
      const int32_t k = c % n_chains;
      g.vs = cfg.prior_vs;  // start at the prior mean, hypo_tremor_mcmc.f90:175-185
      g.qs = cfg.prior_qs;
      g.tc.assign(S, cfg.prior_t_corr);
      g.ac.assign(S, cfg.prior_a_corr);
      for (int32_t j = 0; j < S; ++j) {
        uint32_t w[4];
        Philox::gen(cfg.seed, static_cast<uint32_t>(j), chain_offset + static_cast<uint32_t>(c), PHX_INIT, 1u, w);
        if (cfg.solve_t_corr) g.tc[j] = cfg.prior_t_corr + gauss(w[0], w[1]) * cfg.prior_width_t_corr;
        if (cfg.solve_a_corr) g.ac[j] = cfg.prior_a_corr + gauss(w[2], w[3]) * cfg.prior_width_a_corr;
      }
      if (k < cfg.n_cool) {
        g.temp = 1.0;
      } else if (cfg.ladder == HTM_LADDER_GEOMETRIC) {
        g.temp = std::exp(std::log(cfg.temp_high) * static_cast<double>(k - cfg.n_cool + 1) /
                          static_cast<double>(n_chains - cfg.n_cool));
      } else {
        uint32_t w[4];
        Philox::gen(cfg.seed, 0u, chain_offset + static_cast<uint32_t>(c), PHX_TEMP, 1u, w);
        g.temp = std::exp((Philox::u_co(w[0]) * (1.0 - kEps) + kEps) * std::log(cfg.temp_high));
      }
      g.x.resize(E);
      g.y.resize(E);
      g.z.resize(E);
      g.Le.resize(E);
      g.L = 0.0;
      for (int32_t e = 0; e < E; ++e) {
        uint32_t a[4], b[4];
        Philox::gen(cfg.seed, 0u, gidC(e, c), PHX_INIT, 0u, a);
        Philox::gen(cfg.seed, 1u, gidC(e, c), PHX_INIT, 0u, b);
        g.x[e] = x_mu[e] + gauss(a[0], a[1]) * cfg.prior_width_xy;
        g.y[e] = y_mu[e] + gauss(a[2], a[3]) * cfg.prior_width_xy;
        g.z[e] = cfg.prior_z + std::sqrt(-2.0 * std::log(Philox::u_oo(b[0]))) * cfg.prior_width_z;
        const double xyz[3] = {g.x[e], g.y[e], g.z[e]};
        g.Le[e] = fwd.event_log_likelihood(e, xyz, g.tc.data(), g.vs, g.ac.data(), g.qs);
      }
      g.L = sum_events(g.Le);
    }
  }
  // fixed-order pairwise sum (the device reduces 32-event tiles with a butterfly and adds the
  // tile sums in order; any fixed order is fine at 1e-9, this one is simply left to right)
  static double sum_events(const std::vector<double>& v) {
    double s = 0.0;
    for (double t : v) s += t;
    return s;
  }
  void hypo_step_gibbs(int32_t c, int32_t e, int32_t it, htm_step_trace* tr) {
    GibbsChain& g = gc[c];
    uint32_t w[4];
    Philox::gen(cfg.seed, static_cast<uint32_t>(it), gidC(e, c), PHX_STEP, 0u, w);
    const int32_t icmp = static_cast<int32_t>(Philox::below(w[0], 3));
    const int32_t comp = 2 - icmp;
    const double gs = gauss(w[1], w[2]);
    double xyz[3] = {g.x[e], g.y[e], g.z[e]};
    const double x_old = xyz[comp];
    const double mu = comp == 0 ? x_mu[e] : (comp == 1 ? y_mu[e] : cfg.prior_z);
    const double sigma = comp == 2 ? cfg.prior_width_z : cfg.prior_width_xy;
    const double step = comp == 2 ? cfg.step_size_z : cfg.step_size_xy;
    const double x_new = x_old + gs * step;
    double lpr = -((x_new - mu) * (x_new - mu) - (x_old - mu) * (x_old - mu)) / (2.0 * sigma * sigma);
    bool prior_ok = true;
    if (comp == 2) {
      if (x_new <= mu)
        prior_ok = false;
      else
        lpr = lpr + std::log(x_new - mu) - std::log(x_old - mu);
    }
    const int32_t type = 5 + icmp;
    const bool cold = g.temp < 1.0 + kEps;
    if (cold) b_propose[type - 1] += 1;
    bool acc = false;
    double ll_new = g.Le[e];
    if (prior_ok) {
      xyz[comp] = x_new;
      ll_new = fwd.event_log_likelihood(e, xyz, g.tc.data(), g.vs, g.ac.data(), g.qs);
      const double ratio = (ll_new - g.Le[e]) / g.temp + lpr;
      const double rr = Philox::u_co(w[3]);
      if (rr >= kEps && std::log(rr) <= ratio) acc = true;
    }
    if (acc) {
      g.x[e] = xyz[0];
      g.y[e] = xyz[1];
      g.z[e] = xyz[2];
      g.Le[e] = ll_new;
      if (cold) b_accept[type - 1] += 1;
    }
    if (tr) {
      tr->proposal_type = type;
      tr->index = 3 * (e + 1) - icmp;
      tr->prior_ok = prior_ok ? 1 : 0;
      tr->accepted = acc ? 1 : 0;
      tr->log_likelihood = g.Le[e];
    }
  }
  // Event shards of ONE joint ensemble (cfg.gibbs_shard_events on the device): this oracle holds a slice of
  // the events of every chain; the two sums over events that judge a shared-parameter move must then cover
  // all shards.  The hook receives this shard's {sum of L_e, sum of L_e under the proposal} of one chain and
  // returns the sums over all shards (the host test all-gathers them and adds in shard order, as the device's
  // peer-memory exchange does).  Null = single shard.
  typedef void (*SumHook)(double* two, void* user);
  SumHook sum_hook = nullptr;
  void* sum_hook_user = nullptr;
  void global_step_gibbs(int32_t c, int32_t it, htm_step_trace* tr) {
    GibbsChain& g = gc[c];
    g.L = sum_events(g.Le);
    int32_t types[4], n_g;
    solved_types(types, n_g);
    if (n_g == 0) {
      if (tr) {
        tr->proposal_type = 0;
        tr->index = 0;
        tr->prior_ok = 1;
        tr->accepted = 0;
        tr->log_likelihood = g.L;
      }
      return;
    }
    uint32_t wa[4], wb[4];
    Philox::gen(cfg.seed, static_cast<uint32_t>(it), chain_offset + static_cast<uint32_t>(c), PHX_GLOBAL, 0u, wa);
    Philox::gen(cfg.seed, static_cast<uint32_t>(it), chain_offset + static_cast<uint32_t>(c), PHX_GLOBAL, 1u, wb);
    const int32_t type = types[Philox::below(wa[0], static_cast<uint32_t>(n_g))];
    const int32_t idx = (type == 2 || type == 4) ? static_cast<int32_t>(Philox::below(wa[1], static_cast<uint32_t>(S))) : 0;
    const double gs = gauss(wa[2], wa[3]);
    double* slot;
    double mu, sigma, step;
    if (type == 1) {
      slot = &g.vs; mu = cfg.prior_vs; sigma = cfg.prior_width_vs; step = cfg.step_size_vs;
    } else if (type == 2) {
      slot = &g.tc[idx]; mu = cfg.prior_t_corr; sigma = cfg.prior_width_t_corr; step = cfg.step_size_t_corr;
    } else if (type == 3) {
      slot = &g.qs; mu = cfg.prior_qs; sigma = cfg.prior_width_qs; step = cfg.step_size_qs;
    } else {
      slot = &g.ac[idx]; mu = cfg.prior_a_corr; sigma = cfg.prior_width_a_corr; step = cfg.step_size_a_corr;
    }
    const double x_old = *slot;
    const double x_new = x_old + gs * step;
    const double lpr = -((x_new - mu) * (x_new - mu) - (x_old - mu) * (x_old - mu)) / (2.0 * sigma * sigma);
    const bool cold = g.temp < 1.0 + kEps;
    if (cold) b_propose[type - 1] += 1;
    *slot = x_new;
    std::vector<double> Lp(E);
    for (int32_t e = 0; e < E; ++e) {
      const double xyz[3] = {g.x[e], g.y[e], g.z[e]};
      Lp[e] = fwd.event_log_likelihood(e, xyz, g.tc.data(), g.vs, g.ac.data(), g.qs);
    }
    double L_new = sum_events(Lp);
    if (sum_hook) {
      double two[2] = {g.L, L_new};
      sum_hook(two, sum_hook_user);
      g.L = two[0];
      L_new = two[1];
    }
    const double ratio = (L_new - g.L) / g.temp + lpr;
    const double rr = Philox::u_co(wb[0]);
    bool acc = false;
    if (rr >= kEps && std::log(rr) <= ratio) acc = true;
    if (acc) {
      g.Le = Lp;
      g.L = L_new;
      if (cold) b_accept[type - 1] += 1;
    } else {
      *slot = x_old;
    }
    if (tr) {
      tr->proposal_type = type;
      tr->index = idx + 1;
      tr->prior_ok = 1;
      tr->accepted = acc ? 1 : 0;
      tr->log_likelihood = g.L;
    }
  }
  void record_gibbs(int32_t it) {
    if ((it % cfg.n_interval) != 1) return;
    for (int32_t c = 0; c < J(); ++c) {
      const GibbsChain& g = gc[c];
      if (!(g.temp < 1.0 + kEps)) continue;
      RankOutput& o = out[c / n_chains];
      if (it > cfg.n_burn) {
        o.iter.push_back(it);
        o.vs.push_back(g.vs);
        o.qs.push_back(g.qs);
        for (int32_t e = 0; e < E; ++e) {
          o.hypo.push_back(g.x[e]);
          o.hypo.push_back(g.y[e]);
          o.hypo.push_back(g.z[e]);
        }
        o.t_corr.insert(o.t_corr.end(), g.tc.begin(), g.tc.end());
        o.a_corr.insert(o.a_corr.end(), g.ac.begin(), g.ac.end());
      }
      o.lik_iter.push_back(it);
      o.lik.push_back(g.L);
    }
  }
  void swap_gibbs(int32_t it, htm_swap_trace* tr) {
    const int32_t n = J();
    if (n < 2) return;
    uint32_t w[4];
    Philox::gen(cfg.seed, static_cast<uint32_t>(it), swap_stream, PHX_SWAP, 1u, w);
    const int32_t i1 = static_cast<int32_t>(Philox::below(w[0], static_cast<uint32_t>(n)));
    const int32_t i2 = (i1 + 1 + static_cast<int32_t>(Philox::below(w[1], static_cast<uint32_t>(n - 1)))) % n;
    const bool acc = judge_swap_with(gc[i1].temp, gc[i2].temp, gc[i1].L, gc[i2].L, Philox::u_co(w[2]));
    if (acc) std::swap(gc[i1].temp, gc[i2].temp);
    if (tr) {
      tr->rank1 = i1 / n_chains;
      tr->chain1 = i1 % n_chains + 1;
      tr->rank2 = i2 / n_chains;
      tr->chain2 = i2 % n_chains + 1;
      tr->accepted = acc ? 1 : 0;
      tr->reserved = 0;
    }
  }
  // trace: [iteration][E + 1][J] (row E = the shared-parameter step); swaps: [iteration]
  void run_gibbs(int32_t iter_first, int32_t iter_last, htm_step_trace* trace, htm_swap_trace* swaps) {
    const size_t per_it = static_cast<size_t>(E + 1) * J();
    for (int32_t it = iter_first; it <= iter_last; ++it) {
      htm_step_trace* t = trace ? trace + static_cast<size_t>(it - iter_first) * per_it : nullptr;
      for (int32_t c = 0; c < J(); ++c) {
        for (int32_t e = 0; e < E; ++e) hypo_step_gibbs(c, e, it, t ? t + static_cast<size_t>(e) * J() + c : nullptr);
        global_step_gibbs(c, it, t ? t + static_cast<size_t>(E) * J() + c : nullptr);
      }
      record_gibbs(it);
      swap_gibbs(it, swaps ? swaps + (it - iter_first) : nullptr);
    }
  }
};

}  // namespace hto
