/*
 * htm_b200.h -- C ABI of libhtm_b200.so, the B200 (sm_100a) implementation of the
 * hypo_tremor_mcmc inversion hot path of akuhara/HypoTremorMCMC.
 *
 * The reference has no FFI: its de-facto boundary is the set of Fortran type-bound
 * calls the driver makes in src/hypo_tremor_mcmc.f90:101-118,188-209,236-291.  Those
 * per-proposal calls are far too fine-grained to cross to a GPU, so this ABI moves the
 * WHOLE loop (src/hypo_tremor_mcmc.f90:236-284) behind one call and keeps everything
 * the Fortran driver does around it (parameter file, observation files, output files).
 * Each entry point cites the reference interface it replaces.
 *
 * Conventions (chosen so that a Fortran `bind(C)` interface is mechanical):
 *   - scalars by value, arrays as pointer + extents taken from the handle's config;
 *   - arrays are column-major exactly as Fortran stores them, e.g. t_obs(n_sta,n_events)
 *     is passed as double[n_events][n_sta] (station index fastest);
 *   - only int32_t / int64_t / uint64_t / double cross the boundary; logicals are int32_t;
 *   - the handle is an opaque pointer (type(c_ptr));
 *   - every call returns int32_t, 0 = HTM_OK; htm_last_error() gives the message.
 *     (The reference prints and `stop`s instead, src/cls_parallel.f90:113-117; a library
 *     must not terminate its host, so the driver decides.)
 *   - the library copies inputs during the call and never keeps caller pointers;
 *   - one handle = one host thread = one CUDA device; calls on a handle are not concurrent.
 *   - there is NO CPU fallback: every compute entry point needs a CUDA device and fails
 *     with HTM_ERR_CUDA otherwise.
 */
#ifndef HTM_B200_H
#define HTM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HTM_ABI_VERSION 2

/* ---- status codes --------------------------------------------------------------- */
enum {
  HTM_OK = 0,
  HTM_ERR_ARG = 1,       /* bad argument / inconsistent configuration              */
  HTM_ERR_STATE = 2,     /* call order violated (e.g. run before observations set) */
  HTM_ERR_CUDA = 3,      /* CUDA runtime error or no device                        */
  HTM_ERR_DRAWS = 4,     /* replay: the supplied draw stream was exhausted          */
  HTM_ERR_UNSUPPORTED = 5
};

/* ---- sampling schedule ---------------------------------- */
enum {
  /* A: reference-exact joint chain.  One chain holds ALL events plus vs, qs, t_corr,
   *    a_corr; one scalar is perturbed per iteration; one swap attempt per iteration
   *    over all n_procs*n_chains chains (src/hypo_tremor_mcmc.f90:236-284).  Driven by
   *    htm_replay() with the host's mod_random draws.  Float64 only. */
  HTM_MODE_REPLAY = 0,
  /* B: factorised.  Needs solve_* all false: the posterior factorises over events, so
   *    every (event, virtual rank, chain) is an independent tempered Metropolis chain
   *    over (x,y,z) of ONE event; the chains of one (event, rank) form a tempering group
   *    with one swap attempt per iteration.  Same stationary distribution as A. */
  HTM_MODE_FACTORISED = 1,
  /* C: blocked Gibbs.  Any solve_* true: each of the n_procs*n_chains joint chains
   *    updates all its hypocentres in parallel given its globals, then proposes one
   *    global parameter, judged on the sum over events. */
  HTM_MODE_BLOCKED_GIBBS = 2
};

enum { HTM_PRECISION_F64 = 64, HTM_PRECISION_F32 = 32 };

/* temperature ladder of the hot chains (chain index > n_cool within a rank) */
enum {
  HTM_LADDER_RANDOM = 0,    /* reference: log-uniform random in [1,temp_high],
                               src/hypo_tremor_mcmc.f90:202-208 */
  HTM_LADDER_GEOMETRIC = 1  /* fixed: T_k = temp_high^(k/(n_hot)), k=1..n_hot */
};

/* kernel layout of the factorised mode */
enum {
  HTM_KERNEL_AUTO = 0,
  HTM_KERNEL_WARP_PER_CHAIN = 1, /* stations across the 32 lanes, shuffle reductions */
  HTM_KERNEL_LANE_PER_CHAIN = 2  /* one lane per chain, 32 chains of one event per warp */
};

/* ---- configuration ------------------------------------------------------------------
 * Mirrors the getters the driver reads from cls_param (src/hypo_tremor_mcmc.f90:92-208,
 * src/cls_param.f90 "mcmc" key list :127-137) plus the B200-side switches.
 * Layout: all 8-byte members first, then 4-byte members, so that the Fortran bind(C)
 * derived type in fortran/htm_b200_binding.f90 has no hidden padding questions. */
typedef struct htm_config {
  /* -- 8-byte members -- */
  uint64_t seed;            /* Philox key (modes B, C); ignored by replay               */
  double temp_high;         /* cls_param get_temp_high                                  */
  double prior_z, prior_width_z, prior_width_xy;
  double prior_vs, prior_width_vs, prior_qs, prior_width_qs;
  double prior_t_corr, prior_width_t_corr, prior_a_corr, prior_width_a_corr;
  double step_size_z, step_size_xy, step_size_vs, step_size_qs;
  double step_size_t_corr, step_size_a_corr;
  double hist_xy_halfwidth; /* posterior histograms: x,y bins span prior mean +- this   */
  double hist_z_max;        /* z bins span [prior_z, prior_z + hist_z_max]              */
  /* -- 4-byte members -- */
  int32_t abi_version;      /* must be HTM_ABI_VERSION                                  */
  int32_t n_sta;            /* get_n_stations                                           */
  int32_t n_events;         /* size(win_id), src/hypo_tremor_mcmc.f90:86                */
  int32_t n_procs;          /* virtual MPI ranks (mpi_comm_size in the reference)       */
  int32_t n_chains;         /* chains per rank                                          */
  int32_t n_cool;           /* T=1 chains per rank                                      */
  int32_t n_iter, n_burn, n_interval;
  int32_t solve_vs, solve_t_corr, solve_qs, solve_a_corr;
  int32_t use_time, use_amp;
  int32_t mode;             /* HTM_MODE_*                                               */
  int32_t precision;        /* HTM_PRECISION_*                                          */
  int32_t ladder;           /* HTM_LADDER_*                                             */
  int32_t kernel;           /* HTM_KERNEL_*                                             */
  int32_t device;           /* CUDA device ordinal                                      */
  int32_t shard_rank;       /* this process's index among shard_count shards            */
  int32_t shard_count;      /* modes A/B: EVENTS are split in contiguous blocks over the
                               shards.  Mode C (joint chains): the VIRTUAL RANKS are split
                               instead (every shard holds all events and runs an independent
                               ensemble of its ranks; `rank` arguments are then shard-local),
                               unless gibbs_shard_events = 1 */
  int32_t hist_bins;        /* bins per coordinate histogram; 0 = no histograms         */
  int32_t max_samples;      /* capacity (in recorded iterations) of the sample ring     */
  int32_t lane_slots;       /* lane-per-chain kernel: chains per thread (1, 2, 4); 0 = auto */
  int32_t gibbs_shard_events; /* mode C only: 1 = shard the EVENTS of every joint chain over the shards (all
                               shards hold all chains; one NCCL all-reduce of the per-chain sums per
                               iteration; needs htm_comm_init).  0 = shard the virtual ranks.        */
  int32_t summary;          /* 1 = keep every post-burn-in cold-chain sample in a device-side store for
                               htm_posterior_quantiles (needs max_samples > 0); 0 = off.  (This member also
                               keeps sizeof(htm_config) a multiple of 8.)                                   */
} htm_config;

/* One record per (iteration, rank, chain) step of the replay mode, in loop order
 * (iteration ascending, then rank, then chain).  Fields follow the quantities the
 * reference computes in mcmc_propose_model / mcmc_judge_model (src/cls_mcmc.f90:115-226). */
typedef struct htm_step_trace {
  int32_t proposal_type;  /* i_proposal_type 1..7: vs,t_corr,qs,a_corr, 5+icmp        */
  int32_t index;          /* 1-based perturbed index inside its model (station id, or
                             3*id-icmp for hypocentres; 1 for vs/qs)                  */
  int32_t prior_ok;       /* 0/1                                                      */
  int32_t accepted;       /* 0/1                                                      */
  double log_likelihood;  /* chain's log-likelihood after the judge                   */
} htm_step_trace;

/* One record per iteration: the swap attempt of parallel_swap_temperature
 * (src/cls_parallel.f90:100-216). chain ids are 1-based as in the reference. */
typedef struct htm_swap_trace {
  int32_t rank1, chain1, rank2, chain2;
  int32_t accepted;
  int32_t reserved;
} htm_swap_trace;

typedef struct htm_handle_s* htm_handle;

/* ---- life cycle ---------------------------------------------------------------------- */

/* Fill *cfg with the defaults of sample/hypo_tremor.in:138-267 (priors, step sizes,
 * temp_high, n_chains=5, n_cool=1) and the B200-side defaults.  Sizes are left 0. */
int32_t htm_config_default(htm_config* cfg);

/* Replaces the param getters read at src/hypo_tremor_mcmc.f90:92-208 and the
 * `parallel(...)` / `mcmc(...)` constructors (src/cls_parallel.f90:35-62,
 * src/cls_mcmc.f90:64-111).  Validates the configuration, selects the device. */
int32_t htm_create(htm_handle* out, const htm_config* cfg);
int32_t htm_destroy(htm_handle h);

/* Message of the last failing call on this handle (h may be NULL for create errors).
 * Copies at most len-1 characters and NUL-terminates. */
int32_t htm_last_error(htm_handle h, char* buf, int32_t len);

/* ---- inputs -------------------------------------------------------------------------- */

/* Replaces the sta_x/sta_y/sta_z arguments of `forward(...)`, src/cls_forward.f90:40-55. */
int32_t htm_set_stations(htm_handle h, const double* sta_x, const double* sta_y,
                         const double* sta_z);

/* Replaces obs%get_t_obs() ... obs%get_a_stdv() consumed by init_forward,
 * src/cls_forward.f90:71-92.  Arrays are (n_sta, n_events) column-major for the events
 * OF THIS SHARD.  The library applies the degenerate-sigma rule of :78-90 itself. */
int32_t htm_set_observations(htm_handle h, const double* t_obs, const double* t_stdv,
                             const double* a_obs, const double* a_stdv);

/* Replaces the result of obs%make_initial_guess, src/cls_obs_data.f90:120-134, used as
 * prior mean of x and y at src/hypo_tremor_mcmc.f90:163-164.  [n_events of this shard] */
int32_t htm_set_xy_prior(htm_handle h, const double* x_mu, const double* y_mu);

/* Fixed (not solved) global parameters for mode B; defaults are the prior means, as the
 * reference does for unsolved parameters (src/hypo_tremor_mcmc.f90:133-137,175-185). */
int32_t htm_set_globals(htm_handle h, double vs, double qs, const double* t_corr,
                        const double* a_corr);

/* ---- chain state ----------------------------------------------------------------------- */

/* Replaces the chain set-up loop src/hypo_tremor_mcmc.f90:120-211 with device-side
 * Philox draws (modes B and C): generate_model for hypocentres (and t_corr/a_corr when
 * solved), vs/qs at their prior means, temperatures per cfg.ladder, and the initial
 * log-likelihood of every chain. */
int32_t htm_init_chains(htm_handle h);

/* Set / get one chain's complete state.  `rank` 0-based, `chain` 0-based.
 * Mode A/C: hypo[3*n_events] interleaved x,y,z per event as model%x of the hypo model
 * (src/hypo_tremor_mcmc.f90:162-170); t_corr[n_sta], a_corr[n_sta].
 * Mode B: same layout; globals are ignored on set and returned as the fixed values.
 * log_likelihood: mode A the joint L (may be the -9e300 sentinel of cls_mcmc.f90:88). */
int32_t htm_set_chain_state(htm_handle h, int32_t rank, int32_t chain, const double* hypo,
                            const double* t_corr, const double* a_corr, double vs,
                            double qs, double temp, double log_likelihood);
int32_t htm_get_chain_state(htm_handle h, int32_t rank, int32_t chain, double* hypo,
                            double* t_corr, double* a_corr, double* vs, double* qs,
                            double* temp, double* log_likelihood);

/* ---- the hot path ---------------------------------------------------------------------- */

/* Batched forward likelihood.  Replaces forward%calc_log_likelihood,
 * src/cls_forward.f90:268-303, for n_models models at once:
 * hypo[n_models][3*n_events], t_corr[n_models][n_sta], a_corr[n_models][n_sta],
 * vs[n_models], qs[n_models] -> log_likelihood[n_models].
 * per_event (may be NULL): [n_models][n_events] per-event contributions. */
int32_t htm_loglik(htm_handle h, int32_t n_models, const double* hypo, const double* t_corr,
                   const double* a_corr, const double* vs, const double* qs,
                   double* log_likelihood, double* per_event);

/* Replaces the loop src/hypo_tremor_mcmc.f90:236-284 (propose, forward, judge, record,
 * swap) for iterations iter_first..iter_last (1-based, inclusive), modes B and C, with
 * Philox draws.  Asynchronous: returns after the launches are queued. */
int32_t htm_run(htm_handle h, int32_t iter_first, int32_t iter_last);

/* Validation entry point: htm_run that also returns, for every step, what
 * mcmc_propose_model / mcmc_judge_model decided (trace: [n_it][n_events][n_procs][n_chains])
 * and every swap attempt (swaps: [n_it][n_events][n_procs]), so the kernels can be compared
 * step by step with a CPU statement of the same schedule.  Synchronous; small cases only. */
int32_t htm_run_traced(htm_handle h, int32_t iter_first, int32_t iter_last,
                       htm_step_trace* trace, htm_swap_trace* swaps);

/* Validation entry points of the float32 blocked-Gibbs kernel (no counterpart in the reference, like
 * htm_run_traced).  htm_gibbs_pending: the shared-parameter proposal (mcmc_propose_model's vs / t_corr / qs /
 * a_corr branches, src/cls_mcmc.f90:139-157) that the NEXT iteration will judge, per joint chain of this shard:
 * which = 1 vs, 2 t_corr, 3 qs, 4 a_corr; idx = 0-based station; x_new.  htm_gibbs_last_sums: per joint chain,
 * the sums over this shard's events of the per-event log-likelihoods the LAST iteration judged, current and
 * under its proposal (what forward%calc_log_likelihood returns for the two models, src/cls_forward.f90:268-303;
 * the kernel forms the difference from per-event moments, see csrc/htm_gibbs_f32.cu). */
int32_t htm_gibbs_pending(htm_handle h, int32_t* which, int32_t* idx, double* x_new);
int32_t htm_gibbs_last_sums(htm_handle h, double* cur, double* prop);

/* Wait for queued work; returns the first asynchronous error if any. */
int32_t htm_synchronize(htm_handle h);

/* Mode A: same loop, consuming the host's mod_random output instead of Philox.
 * draws[rank] points at n_draws[rank] raw xorshift128 words (the value of `w` after each
 * update, src/mod_random.f90:63-71) of that virtual rank's stream, starting at the draw
 * that follows whatever was consumed before iter_first.  The GPU converts them exactly
 * as rand_u / rand_u2 do (:72,:90).  trace: [(iter_last-iter_first+1)*n_procs*n_chains],
 * swaps: [iter_last-iter_first+1]; either may be NULL.  n_used[rank] (may be NULL)
 * receives how many words of each stream were consumed. */
int32_t htm_replay(htm_handle h, int32_t iter_first, int32_t iter_last,
                   const int32_t* const* draws, const int64_t* n_draws,
                   htm_step_trace* trace, htm_swap_trace* swaps, int64_t* n_used);

/* ---- outputs --------------------------------------------------------------------------- */

/* Replaces mc%write_out_vs/hypo/t_corr/qs/a_corr + the `write(io_*)` statements at
 * src/hypo_tremor_mcmc.f90:270-278: the thinned cold-chain samples of virtual rank
 * `rank`, in loop order (iteration ascending, then chain), that were recorded since the
 * last fetch.  The driver does the Fortran `write`, so -fconvert stays its business.
 * Any output pointer may be NULL.  hypo: [n][3*n_events(shard)], t_corr/a_corr: [n][n_sta]. */
int32_t htm_fetch_samples(htm_handle h, int32_t rank, int32_t max_records,
                          int32_t* n_records, int32_t* iter, double* vs, double* qs,
                          double* hypo, double* t_corr, double* a_corr);

/* Replaces `write(io_likelihood) i, mc%get_log_likelihood()`, :279 (includes burn-in). */
int32_t htm_fetch_likelihood(htm_handle h, int32_t rank, int32_t max_records,
                             int32_t* n_records, int32_t* iter, double* log_likelihood);

/* Drop every pending sample / likelihood record (callers that only read histograms). */
int32_t htm_discard_samples(htm_handle h);

/* Replaces parallel%output_proposal's reduction, src/cls_parallel.f90:244-268: proposal and
 * acceptance counts of T=1 chains, summed over ranks and chains of this shard.
 * Order as mcmc%label: vs, t_corr, qs, a_corr, "x", "y", "z" (with the reference's
 * icmp quirk: slot 5 counts z moves, slot 7 counts x moves). */
int32_t htm_get_counts(htm_handle h, int64_t n_propose[7], int64_t n_accept[7]);

/* New: device-side posterior histograms of cold-chain hypocentres after burn-in.
 * hist: [n_events(shard)][3][hist_bins] counts (x, y, z). */
int32_t htm_get_histograms(htm_handle h, uint32_t* hist);

/* Device-side form of the tables `hypo_tremor_statistics` computes from the .out files: per marginal the sorted
 * post-burn-in cold-chain sample at the 1-based positions im = 0.5 n, il = 0.025 n, iu = 0.975 n
 * (src/cls_statistics.f90:230-232,360-362; default-real products truncated to integer), in the column order of
 * hypo.stat / station_corrections.stat / uniform_structure.stat: {50 %, 2.5 %, 97.5 %}
 * (src/cls_statistics.f90:216-264,345-431).  Needs cfg.summary = 1; covers everything recorded since
 * htm_init_chains.  n_samples = n_mod of src/cls_statistics.f90:65 once the run is complete.
 * hypo_q: [3*n_events(shard)][3] (x, y, z of event 1, then event 2, ...); vs_q, qs_q: [3];
 * t_corr_q, a_corr_q: [n_sta][3].  Any output pointer may be NULL. */
int32_t htm_posterior_quantiles(htm_handle h, int32_t* n_samples, double* hypo_q, double* vs_q, double* qs_q,
                                double* t_corr_q, double* a_corr_q);

/* ---- multi-GPU (one process per GPU, events sharded by cfg.shard_rank / cfg.shard_count) -----
 * The data path needs no exchange; these calls are the once-per-flush gathers.  NCCL is loaded
 * with dlopen at the first call.  htm_comm_unique_id: called on ONE shard; the HOST program hands
 * the 128 bytes to every shard (MPI_Bcast in a Fortran/MPI driver, a file, torch.distributed ...).
 * htm_gather: all-gather of the event-sharded histograms into hist_all[n_events_total][3][bins]
 * (host, may be NULL) and all-reduce of the proposal counters -- the latter replaces the MPI_Reduce
 * of parallel%output_proposal, src/cls_parallel.f90:265-268.  Collective: every shard must call. */
int32_t htm_comm_unique_id(char id[128]);
int32_t htm_comm_init(htm_handle h, const char id[128]);
int32_t htm_gather(htm_handle h, uint32_t* hist_all, int64_t n_propose[7], int64_t n_accept[7]);

/* Collective form of htm_fetch_samples for the event-sharded factorised mode: every shard receives the records
 * of virtual rank `rank` with the hypocentres of ALL events, hypo_all: [n][3*n_events_total] (all-gather of the
 * shards' blocks over NVLink), so that one process can write the reference's hypo.RR.out records
 * (src/hypo_tremor_mcmc.f90:272-274 writes all 3E values per record).  Every shard must call with the same
 * arguments; it consumes the pending records exactly as htm_fetch_samples does. */
int32_t htm_gather_samples(htm_handle h, int32_t rank, int32_t max_records, int32_t* n_records, int32_t* iter,
                           double* vs, double* qs, double* hypo_all, double* t_corr, double* a_corr);

/* Event-sharded blocked Gibbs (cfg.gibbs_shard_events): the one exchange step of the path -- the sum over
 * ALL events of every joint chain's log-likelihood, needed for each shared-parameter move (the reference
 * recomputes it with forward%calc_log_likelihood, src/cls_forward.f90:268-303, on one rank) -- done inside
 * the sweep kernel through NVLink peer memory instead of a separate all-reduce.  Every shard calls
 * htm_comm_p2p_export, the HOST program all-gathers the 64-byte handles in shard order (MPI_Allgather in a
 * Fortran/MPI driver), every shard calls htm_comm_p2p_import with handles[shard_count][64].  All shards must
 * be processes on one NVLink/NVSwitch box (CUDA IPC).  With both this and htm_comm_init set up, the
 * peer-memory path is used. */
int32_t htm_comm_p2p_export(htm_handle h, unsigned char handle[64]);
int32_t htm_comm_p2p_import(htm_handle h, const unsigned char* handles);

/* Device pointers for zero-copy consumers in the same process (e.g. a NCCL gather driven
 * by the host program).  what: 0 = histograms (uint32), 1 = counts (int64[14]). */
int32_t htm_device_ptr(htm_handle h, int32_t what, void** ptr, int64_t* n_bytes);

/* Timing of the last htm_run on this handle's stream, CUDA events (milliseconds), and the
 * number of kernels it launched. */
int32_t htm_last_run_stats(htm_handle h, double* ms, int64_t* n_launches,
                           int64_t* n_proposals);

/* ---- upstream stage ---------------------------------------------------------------
 * Replaces the loop `do i = i1, i2; call slct%eval_wave_propagation(win_id(i), vs(i), t0(i), b(i), a0(i), cc_t(i),
 * cc_a(i))` of src/hypo_tremor_select.f90:93-96 and the acceptance test of :122-127, for all detected windows at
 * once: per window the station of maximum log-amplitude is taken as the epicentre guess, distances use z_guess
 * (src/cls_selector.f90:61-67), amplitudes are corrected for geometrical spreading, arrival times and amplitudes are
 * fitted against distance with weights 1/err^2 (src/mod_regress.f90:5-38: vs = 1/slope, B = -slope) and the
 * correlation coefficients of weighted_corr (:40-58) are formed.  t, t_err, a, a_err: [n_events][n_sta] as read from
 * opt_data.NNNNNN.dat columns 4-7 (src/cls_selector.f90:92-95).  Outputs [n_events]; selected = 1 where
 * vs_min <= vs <= vs_max and b_min <= B <= b_max (what the driver then writes to selected_win.dat, and all six
 * numbers to regress.dat).  Stand-alone: no handle; float64 throughout; kernel_ms (may be NULL) receives the
 * CUDA-event time of the kernel. */
int32_t htm_select_events(int32_t device, int32_t n_sta, int32_t n_events, const double* sta_x, const double* sta_y,
                          const double* sta_z, double z_guess, const double* t, const double* t_err, const double* a,
                          const double* a_err, double vs_min, double vs_max, double b_min, double b_max, double* vs,
                          double* t0, double* b, double* a0, double* cc_t, double* cc_a, int32_t* selected,
                          double* kernel_ms);

/* Replaces, for all detected windows at once, the numerical core of `call msr%measure_lag_time()`
 * (src/hypo_tremor_measure.f90:57; src/cls_measurer.f90:317-400): per window `optimize_cc` (:466-523 -- 5 % cosine taper,
 * division by the window's sum of squares, circular cross-correlation of every station pair, first maximum -> signed
 * lag, t_i = -(1/S) sum_j lag(i, j) and its scatter) and `optimize_amp` (:405-462 -- envelopes shifted by nint(t_i / dt)
 * samples, log(sxy / sxx_i) per pair, the same averaging and scatter; a negative cross product zeroes the window).  The
 * reference evaluates the correlation with FFTW; the kernel sums it directly in float64 (htm_measure.cu).
 * env: [n_sta][n_total] merged envelopes (the second column of STA.merged.env), sampling interval dt; window w covers
 * samples (win_id[w] - 1) * n_step ... + n_smp - 1 (0-based; win_id as in detected_win.dat, n_smp = nint(t_win / dt),
 * n_step = nint(t_step / dt)).  Outputs [n_win][n_sta] = columns 4-7 of opt_data.NNNNNN.dat (:388-397); lag (may be
 * NULL) [n_win][n_sta (n_sta - 1) / 2] receives each pair's 0-based arg-max sample, pairs in the order i < j of :488-489.
 * Stand-alone: no handle; kernel_ms (may be NULL) receives the CUDA-event time of the kernel.  HTM_ERR_UNSUPPORTED when
 * a window does not fit one CTA's shared memory (about n_sta (n_smp + 48) 8 B + n_sta^2 8 B <= 227 KB: 50 stations x 300
 * samples use 176 KB; 20 stations reach 1 300 samples). */
int32_t htm_measure_windows(int32_t device, int32_t n_sta, int64_t n_total, const double* env, double dt, int32_t n_smp,
                            int32_t n_step, int32_t n_win, const int32_t* win_id, double* t, double* t_stdv, double* amp,
                            double* amp_stdv, int32_t* lag, double* kernel_ms);

/* Replaces `call msr%scan_cc()` (src/hypo_tremor_measure.f90:55; src/cls_measurer.f90:189-315) TOGETHER WITH the
 * correlation functions it reads: scan_cc takes, per station pair, the order statistic int(n n_win alpha) of ALL
 * correlation values of all windows as the pair's threshold (:213-223), marks the windows whose maximum correlation
 * reaches it (:228-236) and detects a window when more than n_pair_thred pairs are marked (:247-253).  The reference
 * reads those values from the STA1.STA2.corr / .max_corr files that hypo_tremor_correlate wrote (24 B per value: 10 GB
 * for two days of 50 stations); here they are recomputed from the envelopes on the device, as run_cross_corr does
 * (src/cls_correlator.f90:200-233: 5 % taper, mean removed, unit length, circular cross-correlation / n), at about
 * 10 us per window, kept in device memory, and the thresholds are selected there (radix selection, no sort).
 * env: [n_sta][n_total] merged envelopes; window w (1-based) = samples (w - 1) n_step ... + n_smp - 1, w = 1 .. n_win with
 * n_win = (n_total - n_smp) / n_step as in src/cls_correlator.f90:80; n_smp even (:180-183).  Outputs: cc_thred [n_pair]
 * (cc_thred.dat), detected [n_win] (1 = listed in detected_win.dat) and, if not NULL, cc_max [n_pair][n_win] (the
 * .max_corr values) and n_pairs_above [n_win]; pairs in the order i < j.  HTM_ERR_UNSUPPORTED when n_pair n_win n_smp
 * 8 B exceeds the free device memory (split the time range) or a window exceeds one CTA (see htm_measure_windows). */
int32_t htm_detect_windows(int32_t device, int32_t n_sta, int64_t n_total, const double* env, int32_t n_smp, int32_t n_step,
                           double alpha, int32_t n_pair_thred, int32_t n_win, double* cc_thred, double* cc_max,
                           int32_t* detected, int32_t* n_pairs_above, double* kernel_ms);

/* FFMA/MUFU microbenchmark used as the FP32 roofline denominator (no driver-measured
 * FP32 vector peak exists).  Returns TFLOP/s (FFMA = 2 flops). */
int32_t htm_measure_fp32_peak(int32_t device, double* tflops, double* mufu_gops);
/* The same for the float64 instantiations (the validation path): DFMA microbenchmark, TFLOP/s. */
int32_t htm_measure_fp64_peak(int32_t device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* HTM_B200_H */
