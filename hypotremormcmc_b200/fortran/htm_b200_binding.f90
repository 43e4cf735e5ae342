!=======================================================================
! htm_b200_binding -- ISO_C_BINDING interface to libhtm_b200.so
! (include/htm_b200.h).  One interface per exported function; the derived
! type mirrors struct htm_config member by member (8-byte members first,
! then 4-byte members, so there is no padding to think about).
!
! NOT compile-tested in the build image (it has no Fortran compiler);
! tests/test_abi.py checks the member list and the bind(C) names against
! the header mechanically.  Fortran 2008, no extensions.
!=======================================================================
module htm_b200_binding
  use, intrinsic :: iso_c_binding
  implicit none

  integer(c_int32_t), parameter :: HTM_ABI_VERSION = 2
  integer(c_int32_t), parameter :: HTM_OK = 0
  integer(c_int32_t), parameter :: HTM_MODE_REPLAY = 0
  integer(c_int32_t), parameter :: HTM_MODE_FACTORISED = 1
  integer(c_int32_t), parameter :: HTM_MODE_BLOCKED_GIBBS = 2
  integer(c_int32_t), parameter :: HTM_PRECISION_F64 = 64
  integer(c_int32_t), parameter :: HTM_PRECISION_F32 = 32
  integer(c_int32_t), parameter :: HTM_LADDER_RANDOM = 0
  integer(c_int32_t), parameter :: HTM_LADDER_GEOMETRIC = 1

  type, bind(c) :: htm_config
     integer(c_int64_t) :: seed
     real(c_double) :: temp_high
     real(c_double) :: prior_z, prior_width_z, prior_width_xy
     real(c_double) :: prior_vs, prior_width_vs, prior_qs, prior_width_qs
     real(c_double) :: prior_t_corr, prior_width_t_corr, prior_a_corr, prior_width_a_corr
     real(c_double) :: step_size_z, step_size_xy, step_size_vs, step_size_qs
     real(c_double) :: step_size_t_corr, step_size_a_corr
     real(c_double) :: hist_xy_halfwidth
     real(c_double) :: hist_z_max
     integer(c_int32_t) :: abi_version
     integer(c_int32_t) :: n_sta
     integer(c_int32_t) :: n_events
     integer(c_int32_t) :: n_procs
     integer(c_int32_t) :: n_chains
     integer(c_int32_t) :: n_cool
     integer(c_int32_t) :: n_iter, n_burn, n_interval
     integer(c_int32_t) :: solve_vs, solve_t_corr, solve_qs, solve_a_corr
     integer(c_int32_t) :: use_time, use_amp
     integer(c_int32_t) :: mode
     integer(c_int32_t) :: precision
     integer(c_int32_t) :: ladder
     integer(c_int32_t) :: kernel
     integer(c_int32_t) :: device
     integer(c_int32_t) :: shard_rank
     integer(c_int32_t) :: shard_count
     integer(c_int32_t) :: hist_bins
     integer(c_int32_t) :: max_samples
     integer(c_int32_t) :: lane_slots
     integer(c_int32_t) :: gibbs_shard_events
     integer(c_int32_t) :: summary
  end type htm_config

  type, bind(c) :: htm_step_trace
     integer(c_int32_t) :: proposal_type, idx, prior_ok, accepted
     real(c_double) :: log_likelihood
  end type htm_step_trace

  type, bind(c) :: htm_swap_trace
     integer(c_int32_t) :: rank1, chain1, rank2, chain2, accepted, reserved
  end type htm_swap_trace

  interface
     function htm_config_default(cfg) bind(c, name="htm_config_default") result(rc)
       import
       type(htm_config), intent(out) :: cfg
       integer(c_int32_t) :: rc
     end function htm_config_default

     function htm_create(h, cfg) bind(c, name="htm_create") result(rc)
       import
       type(c_ptr), intent(out) :: h
       type(htm_config), intent(in) :: cfg
       integer(c_int32_t) :: rc
     end function htm_create

     function htm_destroy(h) bind(c, name="htm_destroy") result(rc)
       import
       type(c_ptr), value :: h
       integer(c_int32_t) :: rc
     end function htm_destroy

     function htm_last_error(h, buf, len) bind(c, name="htm_last_error") result(rc)
       import
       type(c_ptr), value :: h
       character(kind=c_char), intent(out) :: buf(*)
       integer(c_int32_t), value :: len
       integer(c_int32_t) :: rc
     end function htm_last_error

     function htm_set_stations(h, sta_x, sta_y, sta_z) bind(c, name="htm_set_stations") result(rc)
       import
       type(c_ptr), value :: h
       real(c_double), intent(in) :: sta_x(*), sta_y(*), sta_z(*)
       integer(c_int32_t) :: rc
     end function htm_set_stations

     function htm_set_observations(h, t_obs, t_stdv, a_obs, a_stdv) &
          & bind(c, name="htm_set_observations") result(rc)
       import
       type(c_ptr), value :: h
       real(c_double), intent(in) :: t_obs(*), t_stdv(*), a_obs(*), a_stdv(*)
       integer(c_int32_t) :: rc
     end function htm_set_observations

     function htm_set_xy_prior(h, x_mu, y_mu) bind(c, name="htm_set_xy_prior") result(rc)
       import
       type(c_ptr), value :: h
       real(c_double), intent(in) :: x_mu(*), y_mu(*)
       integer(c_int32_t) :: rc
     end function htm_set_xy_prior

     function htm_set_globals(h, vs, qs, t_corr, a_corr) bind(c, name="htm_set_globals") result(rc)
       import
       type(c_ptr), value :: h
       real(c_double), value :: vs, qs
       real(c_double), intent(in) :: t_corr(*), a_corr(*)
       integer(c_int32_t) :: rc
     end function htm_set_globals

     function htm_init_chains(h) bind(c, name="htm_init_chains") result(rc)
       import
       type(c_ptr), value :: h
       integer(c_int32_t) :: rc
     end function htm_init_chains

     function htm_set_chain_state(h, rank, chain, hypo, t_corr, a_corr, vs, qs, temp, &
          & log_likelihood) bind(c, name="htm_set_chain_state") result(rc)
       import
       type(c_ptr), value :: h
       integer(c_int32_t), value :: rank, chain
       real(c_double), intent(in) :: hypo(*), t_corr(*), a_corr(*)
       real(c_double), value :: vs, qs, temp, log_likelihood
       integer(c_int32_t) :: rc
     end function htm_set_chain_state

     function htm_get_chain_state(h, rank, chain, hypo, t_corr, a_corr, vs, qs, temp, &
          & log_likelihood) bind(c, name="htm_get_chain_state") result(rc)
       import
       type(c_ptr), value :: h
       integer(c_int32_t), value :: rank, chain
       real(c_double), intent(out) :: hypo(*), t_corr(*), a_corr(*)
       real(c_double), intent(out) :: vs, qs, temp, log_likelihood
       integer(c_int32_t) :: rc
     end function htm_get_chain_state

     function htm_loglik(h, n_models, hypo, t_corr, a_corr, vs, qs, log_likelihood, per_event) &
          & bind(c, name="htm_loglik") result(rc)
       import
       type(c_ptr), value :: h
       integer(c_int32_t), value :: n_models
       real(c_double), intent(in) :: hypo(*), t_corr(*), a_corr(*), vs(*), qs(*)
       real(c_double), intent(out) :: log_likelihood(*)
       type(c_ptr), value :: per_event
       integer(c_int32_t) :: rc
     end function htm_loglik

     function htm_run(h, iter_first, iter_last) bind(c, name="htm_run") result(rc)
       import
       type(c_ptr), value :: h
       integer(c_int32_t), value :: iter_first, iter_last
       integer(c_int32_t) :: rc
     end function htm_run

     function htm_run_traced(h, iter_first, iter_last, trace, swaps) &
          & bind(c, name="htm_run_traced") result(rc)
       import
       type(c_ptr), value :: h
       integer(c_int32_t), value :: iter_first, iter_last
       type(c_ptr), value :: trace, swaps
       integer(c_int32_t) :: rc
     end function htm_run_traced

     function htm_synchronize(h) bind(c, name="htm_synchronize") result(rc)
       import
       type(c_ptr), value :: h
       integer(c_int32_t) :: rc
     end function htm_synchronize

     function htm_replay(h, iter_first, iter_last, draws, n_draws, trace, swaps, n_used) &
          & bind(c, name="htm_replay") result(rc)
       import
       type(c_ptr), value :: h
       integer(c_int32_t), value :: iter_first, iter_last
       type(c_ptr), intent(in) :: draws(*)
       integer(c_int64_t), intent(in) :: n_draws(*)
       type(c_ptr), value :: trace, swaps
       integer(c_int64_t), intent(out) :: n_used(*)
       integer(c_int32_t) :: rc
     end function htm_replay

     function htm_fetch_samples(h, rank, max_records, n_records, iter, vs, qs, hypo, t_corr, &
          & a_corr) bind(c, name="htm_fetch_samples") result(rc)
       import
       type(c_ptr), value :: h
       integer(c_int32_t), value :: rank, max_records
       integer(c_int32_t), intent(out) :: n_records
       integer(c_int32_t), intent(out) :: iter(*)
       real(c_double), intent(out) :: vs(*), qs(*), hypo(*), t_corr(*), a_corr(*)
       integer(c_int32_t) :: rc
     end function htm_fetch_samples

     function htm_fetch_likelihood(h, rank, max_records, n_records, iter, log_likelihood) &
          & bind(c, name="htm_fetch_likelihood") result(rc)
       import
       type(c_ptr), value :: h
       integer(c_int32_t), value :: rank, max_records
       integer(c_int32_t), intent(out) :: n_records
       integer(c_int32_t), intent(out) :: iter(*)
       real(c_double), intent(out) :: log_likelihood(*)
       integer(c_int32_t) :: rc
     end function htm_fetch_likelihood

     function htm_discard_samples(h) bind(c, name="htm_discard_samples") result(rc)
       import
       type(c_ptr), value :: h
       integer(c_int32_t) :: rc
     end function htm_discard_samples

     function htm_get_counts(h, n_propose, n_accept) bind(c, name="htm_get_counts") result(rc)
       import
       type(c_ptr), value :: h
       integer(c_int64_t), intent(out) :: n_propose(7), n_accept(7)
       integer(c_int32_t) :: rc
     end function htm_get_counts

     function htm_get_histograms(h, hist) bind(c, name="htm_get_histograms") result(rc)
       import
       type(c_ptr), value :: h
       integer(c_int32_t), intent(out) :: hist(*)
       integer(c_int32_t) :: rc
     end function htm_get_histograms

     function htm_device_ptr(h, what, ptr, n_bytes) bind(c, name="htm_device_ptr") result(rc)
       import
       type(c_ptr), value :: h
       integer(c_int32_t), value :: what
       type(c_ptr), intent(out) :: ptr
       integer(c_int64_t), intent(out) :: n_bytes
       integer(c_int32_t) :: rc
     end function htm_device_ptr

     function htm_last_run_stats(h, ms, n_launches, n_proposals) &
          & bind(c, name="htm_last_run_stats") result(rc)
       import
       type(c_ptr), value :: h
       real(c_double), intent(out) :: ms
       integer(c_int64_t), intent(out) :: n_launches, n_proposals
       integer(c_int32_t) :: rc
     end function htm_last_run_stats

     function htm_comm_unique_id(id) bind(c, name="htm_comm_unique_id") result(rc)
       import
       character(kind=c_char), intent(out) :: id(128)
       integer(c_int32_t) :: rc
     end function htm_comm_unique_id

     function htm_comm_init(h, id) bind(c, name="htm_comm_init") result(rc)
       import
       type(c_ptr), value :: h
       character(kind=c_char), intent(in) :: id(128)
       integer(c_int32_t) :: rc
     end function htm_comm_init

     function htm_gather(h, hist_all, n_propose, n_accept) bind(c, name="htm_gather") result(rc)
       import
       type(c_ptr), value :: h
       type(c_ptr), value :: hist_all
       integer(c_int64_t), intent(out) :: n_propose(7), n_accept(7)
       integer(c_int32_t) :: rc
     end function htm_gather

     function htm_gather_samples(h, rank, max_records, n_records, iter, vs, qs, hypo_all, t_corr, a_corr) &
          & bind(c, name="htm_gather_samples") result(rc)
       import
       type(c_ptr), value :: h
       integer(c_int32_t), value :: rank, max_records
       integer(c_int32_t), intent(out) :: n_records
       integer(c_int32_t), intent(out) :: iter(*)
       real(c_double), intent(out) :: vs(*), qs(*), hypo_all(*), t_corr(*), a_corr(*)
       integer(c_int32_t) :: rc
     end function htm_gather_samples

     ! device-side quantile tables of hypo_tremor_statistics (cfg%summary = 1)
     function htm_posterior_quantiles(h, n_samples, hypo_q, vs_q, qs_q, t_corr_q, a_corr_q) &
          & bind(c, name="htm_posterior_quantiles") result(rc)
       import
       type(c_ptr), value :: h
       integer(c_int32_t), intent(out) :: n_samples
       real(c_double), intent(out) :: hypo_q(3,*), vs_q(3), qs_q(3), t_corr_q(3,*), a_corr_q(3,*)
       integer(c_int32_t) :: rc
     end function htm_posterior_quantiles

     ! validation entry points of the float32 blocked-Gibbs kernel (per joint chain of this shard)
     function htm_gibbs_pending(h, which, idx, x_new) bind(c, name="htm_gibbs_pending") result(rc)
       import
       type(c_ptr), value :: h
       integer(c_int32_t), intent(out) :: which(*), idx(*)
       real(c_double), intent(out) :: x_new(*)
       integer(c_int32_t) :: rc
     end function htm_gibbs_pending

     function htm_gibbs_last_sums(h, cur, prop) bind(c, name="htm_gibbs_last_sums") result(rc)
       import
       type(c_ptr), value :: h
       real(c_double), intent(out) :: cur(*), prop(*)
       integer(c_int32_t) :: rc
     end function htm_gibbs_last_sums

     ! handles of all shards, 64 bytes each, in shard order (MPI_Allgather of the exported one)
     function htm_comm_p2p_export(h, handle) bind(c, name="htm_comm_p2p_export") result(rc)
       import
       type(c_ptr), value :: h
       character(kind=c_char), intent(out) :: handle(64)
       integer(c_int32_t) :: rc
     end function htm_comm_p2p_export

     function htm_comm_p2p_import(h, handles) bind(c, name="htm_comm_p2p_import") result(rc)
       import
       type(c_ptr), value :: h
       character(kind=c_char), intent(in) :: handles(*)
       integer(c_int32_t) :: rc
     end function htm_comm_p2p_import

     ! hypo_tremor_select for all windows at once (arrays (n_sta, n_events) column-major)
     function htm_select_events(device, n_sta, n_events, sta_x, sta_y, sta_z, z_guess, t, t_err, a, a_err, &
          & vs_min, vs_max, b_min, b_max, vs, t0, b, a0, cc_t, cc_a, selected, kernel_ms) &
          & bind(c, name="htm_select_events") result(rc)
       import
       integer(c_int32_t), value :: device, n_sta, n_events
       real(c_double), intent(in) :: sta_x(*), sta_y(*), sta_z(*), t(*), t_err(*), a(*), a_err(*)
       real(c_double), value :: z_guess, vs_min, vs_max, b_min, b_max
       real(c_double), intent(out) :: vs(*), t0(*), b(*), a0(*), cc_t(*), cc_a(*)
       integer(c_int32_t), intent(out) :: selected(*)
       real(c_double), intent(out) :: kernel_ms
       integer(c_int32_t) :: rc
     end function htm_select_events

     ! hypo_tremor_measure: optimize_cc + optimize_amp for all detected windows (env (n_total, n_sta) column-major,
     ! outputs (n_sta, n_win)); pass c_null_ptr for lag to skip the per-pair arg-max samples
     function htm_measure_windows(device, n_sta, n_total, env, dt, n_smp, n_step, n_win, win_id, t, t_stdv, amp, &
          & amp_stdv, lag, kernel_ms) bind(c, name="htm_measure_windows") result(rc)
       import
       integer(c_int32_t), value :: device, n_sta, n_smp, n_step, n_win
       integer(c_int64_t), value :: n_total
       real(c_double), intent(in) :: env(*)
       real(c_double), value :: dt
       integer(c_int32_t), intent(in) :: win_id(*)
       real(c_double), intent(out) :: t(*), t_stdv(*), amp(*), amp_stdv(*)
       type(c_ptr), value :: lag
       real(c_double), intent(out) :: kernel_ms
       integer(c_int32_t) :: rc
     end function htm_measure_windows

     ! hypo_tremor_measure: scan_cc with the correlation functions recomputed from the envelopes (env (n_total, n_sta)
     ! column-major; cc_max (n_win, n_pair)); c_null_ptr for cc_max / n_pairs_above to skip them
     function htm_detect_windows(device, n_sta, n_total, env, n_smp, n_step, alpha, n_pair_thred, n_win, cc_thred, &
          & cc_max, detected, n_pairs_above, kernel_ms) bind(c, name="htm_detect_windows") result(rc)
       import
       integer(c_int32_t), value :: device, n_sta, n_smp, n_step, n_pair_thred, n_win
       integer(c_int64_t), value :: n_total
       real(c_double), intent(in) :: env(*)
       real(c_double), value :: alpha
       real(c_double), intent(out) :: cc_thred(*)
       type(c_ptr), value :: cc_max
       integer(c_int32_t), intent(out) :: detected(*)
       type(c_ptr), value :: n_pairs_above
       real(c_double), intent(out) :: kernel_ms
       integer(c_int32_t) :: rc
     end function htm_detect_windows

     function htm_measure_fp64_peak(device, tflops) bind(c, name="htm_measure_fp64_peak") result(rc)
       import
       integer(c_int32_t), value :: device
       real(c_double), intent(out) :: tflops
       integer(c_int32_t) :: rc
     end function htm_measure_fp64_peak

     function htm_measure_fp32_peak(device, tflops, mufu_gops) &
          & bind(c, name="htm_measure_fp32_peak") result(rc)
       import
       integer(c_int32_t), value :: device
       real(c_double), intent(out) :: tflops, mufu_gops
       integer(c_int32_t) :: rc
     end function htm_measure_fp32_peak
  end interface

contains

  ! Stop with the library's message: the reference's error convention is
  ! "print and stop" (e.g. src/cls_parallel.f90:113-117).
  subroutine htm_check(h, rc, where)
    type(c_ptr), intent(in) :: h
    integer(c_int32_t), intent(in) :: rc
    character(*), intent(in) :: where
    character(kind=c_char) :: buf(512)
    character(512) :: msg
    integer :: i
    integer(c_int32_t) :: ignore
    if (rc == HTM_OK) return
    ignore = htm_last_error(h, buf, 512_c_int32_t)
    msg = ""
    do i = 1, 512
       if (buf(i) == c_null_char) exit
       msg(i:i) = buf(i)
    end do
    write(0,*) "ERROR: ", where, ": ", trim(msg)
    error stop
  end subroutine htm_check

end module htm_b200_binding
