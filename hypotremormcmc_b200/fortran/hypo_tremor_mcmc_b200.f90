!=======================================================================
! hypo_tremor_mcmc (B200 edition) -- drop-in for the reference executable
! of the same name.  Same command line (one argument: the parameter
! file), same inputs in the working directory (station file,
! selected_win.dat, opt_data.NNNNNN.dat), same six output file families
! per virtual rank plus proposal_count.txt.
!
! What stays Fortran: parameter parsing (cls_param), observation loading
! (cls_obs_data), every write().  What moved: the loop of the reference
! driver (its lines 236-284) now runs on the GPU behind htm_run(); this
! program only feeds it and drains it in chunks.
!
! MPI is gone: the parameter file's n_procs is the number of VIRTUAL
! ranks (independent stream / file set each), so hypo_tremor_statistics
! run with the same n_procs reads the outputs unchanged.
!
! Build (outside this image, which has no Fortran compiler):
!   gfortran -O2 -fconvert=big-endian cls_line_text.f90 cls_param.f90 \
!       cls_obs_data.f90 htm_b200_binding.f90 hypo_tremor_mcmc_b200.f90 \
!       -L<dir of libhtm_b200.so> -lhtm_b200 -o hypo_tremor_mcmc
! (cls_*.f90 are the reference's own, unmodified sources.)
!=======================================================================
program hypo_tremor_mcmc_b200
  use, intrinsic :: iso_c_binding
  use, intrinsic :: iso_fortran_env, only: iostat_end
  use cls_param, only: param
  use cls_obs_data, only: obs_data
  use cls_line_text, only: line_max
  use htm_b200_binding
  implicit none

  integer, parameter :: chunk_records = 64   ! recorded iterations per GPU launch
  type(param) :: para
  type(obs_data) :: obs
  type(htm_config) :: cfg
  type(c_ptr) :: h
  character(line_max) :: param_file
  character(5), parameter :: label(7) = [character(5) :: "vs", "t_cor", "qs", "a_cor", "x", "y", "z"]
  integer, allocatable :: win_id(:), unit_of(:,:)
  integer :: n_events, n_sta, n_ranks, ios, id, r, k, io
  integer :: it0, it1, n_iter, n_int
  double precision :: t_dummy
  double precision, allocatable :: x_mu(:), y_mu(:)
  logical :: any_solve
  integer(c_int64_t) :: n_prop(7), n_acc(7)

  h = c_null_ptr
  if (command_argument_count() /= 1) error stop "USAGE: hypo_tremor_mcmc [parameter file]"
  call get_command_argument(1, param_file)
  para = param(param_file, verb=.true., from_where="mcmc")

  ! --- events to locate ---
  allocate(win_id(0))
  open(newunit=io, file="selected_win.dat", status="old", iostat=ios)
  if (ios /= 0) error stop "cannot open selected_win.dat"
  do
     read(io, *, iostat=ios) id, t_dummy
     if (ios == iostat_end) exit
     win_id = [win_id, id]
  end do
  close(io)
  n_events = size(win_id)
  n_sta = para%get_n_stations()
  n_ranks = para%get_n_procs()

  obs = obs_data(win_id=win_id, n_sta=n_sta, sta_x=para%get_sta_x(), &
       & sta_y=para%get_sta_y(), verb=.true.)
  allocate(x_mu(n_events), y_mu(n_events))
  call obs%make_initial_guess(x_mu, y_mu)

  ! --- configuration: every number comes from the parameter file ---
  call htm_check(h, htm_config_default(cfg), "htm_config_default")
  cfg%n_sta = n_sta;                 cfg%n_events = n_events
  cfg%n_procs = n_ranks;             cfg%n_chains = para%get_n_chains()
  cfg%n_cool = para%get_n_cool();    cfg%temp_high = para%get_temp_high()
  cfg%n_iter = para%get_n_iter();    cfg%n_burn = para%get_n_burn()
  cfg%n_interval = para%get_n_interval()
  cfg%prior_z = para%get_prior_z();  cfg%prior_width_z = para%get_prior_width_z()
  cfg%prior_width_xy = para%get_prior_width_xy()
  cfg%prior_vs = para%get_prior_vs(); cfg%prior_width_vs = para%get_prior_width_vs()
  cfg%prior_qs = para%get_prior_qs(); cfg%prior_width_qs = para%get_prior_width_qs()
  cfg%prior_t_corr = para%get_prior_t_corr()
  cfg%prior_width_t_corr = para%get_prior_width_t_corr()
  cfg%prior_a_corr = para%get_prior_a_corr()
  cfg%prior_width_a_corr = para%get_prior_width_a_corr()
  cfg%step_size_z = para%get_step_size_z();   cfg%step_size_xy = para%get_step_size_xy()
  cfg%step_size_vs = para%get_step_size_vs(); cfg%step_size_qs = para%get_step_size_qs()
  cfg%step_size_t_corr = para%get_step_size_t_corr()
  cfg%step_size_a_corr = para%get_step_size_a_corr()
  cfg%solve_vs = merge(1, 0, para%get_solve_vs())
  cfg%solve_t_corr = merge(1, 0, para%get_solve_t_corr())
  cfg%solve_qs = merge(1, 0, para%get_solve_qs())
  cfg%solve_a_corr = merge(1, 0, para%get_solve_a_corr())
  cfg%use_time = merge(1, 0, para%get_use_time())
  cfg%use_amp = merge(1, 0, para%get_use_amp())
  any_solve = para%get_solve_vs() .or. para%get_solve_t_corr() .or. &
       & para%get_solve_qs() .or. para%get_solve_a_corr()
  cfg%mode = merge(HTM_MODE_BLOCKED_GIBBS, HTM_MODE_FACTORISED, any_solve)
  cfg%precision = HTM_PRECISION_F32
  cfg%max_samples = chunk_records + 1

  call htm_check(h, htm_create(h, cfg), "htm_create")
  call htm_check(h, htm_set_stations(h, para%get_sta_x(), para%get_sta_y(), para%get_sta_z()), &
       & "htm_set_stations")
  call htm_check(h, htm_set_observations(h, obs%get_t_obs(), obs%get_t_stdv(), &
       & obs%get_a_obs(), obs%get_a_stdv()), "htm_set_observations")
  call htm_check(h, htm_set_xy_prior(h, x_mu, y_mu), "htm_set_xy_prior")
  call htm_check(h, htm_init_chains(h), "htm_init_chains")

  ! --- six stream files per virtual rank, names as the reference writes them ---
  allocate(unit_of(6, 0:n_ranks-1))
  do r = 0, n_ranks - 1
     call open_rank_file(unit_of(1, r), "hypo.", r)
     call open_rank_file(unit_of(2, r), "t_corr.", r)
     call open_rank_file(unit_of(3, r), "vs.", r)
     call open_rank_file(unit_of(4, r), "a_corr.", r)
     call open_rank_file(unit_of(5, r), "qs.", r)
     call open_rank_file(unit_of(6, r), "likelihood", r)   ! sic: no dot (reference file name)
  end do

  ! --- the loop: launch a chunk on the GPU, drain its records, repeat ---
  print *, "start MCMC"
  n_iter = cfg%n_iter
  n_int = max(1, cfg%n_interval)
  it0 = 1
  do while (it0 <= n_iter)
     it1 = min(n_iter, it0 + chunk_records * n_int - 1)
     call htm_check(h, htm_run(h, it0, it1), "htm_run")
     do r = 0, n_ranks - 1
        call drain_rank(r)
     end do
     it0 = it1 + 1
  end do

  call htm_check(h, htm_get_counts(h, n_prop, n_acc), "htm_get_counts")
  open(newunit=io, file="proposal_count.txt", status="unknown")
  do k = 1, 7
     write(io, '(A,2I20)') '"' // label(k) // '"', n_prop(k), n_acc(k)
  end do
  close(io)
  do r = 0, n_ranks - 1
     do k = 1, 6
        close(unit_of(k, r))
     end do
  end do
  call htm_check(h, htm_destroy(h), "htm_destroy")
  stop

contains

  subroutine open_rank_file(u, stem, rank)
    integer, intent(out) :: u
    character(*), intent(in) :: stem
    integer, intent(in) :: rank
    character(line_max) :: fname
    write(fname, '(A,I2.2,A)') stem, rank, ".out"
    open(newunit=u, file=fname, status="replace", access="stream", form="unformatted")
  end subroutine open_rank_file

  ! One write per record: int32 iteration number followed by float64 values, exactly the
  ! record the reference emits.  -fconvert stays the compiler's business.
  subroutine drain_rank(rank)
    integer, intent(in) :: rank
    integer(c_int32_t) :: n, m
    integer(c_int32_t), allocatable :: it(:)
    double precision, allocatable :: vs(:), qs(:), hy(:,:), tc(:,:), ac(:,:), lk(:)
    integer :: cap, j
    cap = (chunk_records + 1) * cfg%n_cool * n_ranks   ! a rank can momentarily hold every cold chain
    allocate(it(cap), vs(cap), qs(cap), hy(3*n_events, cap), tc(n_sta, cap), ac(n_sta, cap), lk(cap))
    call htm_check(h, htm_fetch_samples(h, rank, cap, n, it, vs, qs, hy, tc, ac), "htm_fetch_samples")
    do j = 1, n
       write(unit_of(3, rank)) it(j), vs(j)
       write(unit_of(1, rank)) it(j), hy(:, j)
       write(unit_of(2, rank)) it(j), tc(:, j)
       write(unit_of(5, rank)) it(j), qs(j)
       write(unit_of(4, rank)) it(j), ac(:, j)
    end do
    call htm_check(h, htm_fetch_likelihood(h, rank, cap, m, it, lk), "htm_fetch_likelihood")
    do j = 1, m
       write(unit_of(6, rank)) it(j), lk(j)
    end do
  end subroutine drain_rank

end program hypo_tremor_mcmc_b200
