"""ctypes mirror of ``htm_config`` (include/htm_b200.h) and the sample-file defaults.

The field list follows the getters the reference driver reads from ``cls_param``
(src/hypo_tremor_mcmc.f90:92-208); default values are sample/hypo_tremor.in:138-267.
"""
import ctypes

HTM_ABI_VERSION = 2

HTM_OK, HTM_ERR_ARG, HTM_ERR_STATE, HTM_ERR_CUDA, HTM_ERR_DRAWS, HTM_ERR_UNSUPPORTED = range(6)

MODE_REPLAY, MODE_FACTORISED, MODE_BLOCKED_GIBBS = 0, 1, 2
PRECISION_F64, PRECISION_F32 = 64, 32
LADDER_RANDOM, LADDER_GEOMETRIC = 0, 1
KERNEL_AUTO, KERNEL_WARP_PER_CHAIN, KERNEL_LANE_PER_CHAIN = 0, 1, 2

# mcmc%label, src/cls_mcmc.f90:83 (character(5) truncates "t_corr"/"a_corr")
PROPOSAL_LABELS = ["vs   ", "t_cor", "qs   ", "a_cor", "x    ", "y    ", "z    "]


class HtmConfig(ctypes.Structure):
    _fields_ = [
        ("seed", ctypes.c_uint64),
        ("temp_high", ctypes.c_double),
        ("prior_z", ctypes.c_double),
        ("prior_width_z", ctypes.c_double),
        ("prior_width_xy", ctypes.c_double),
        ("prior_vs", ctypes.c_double),
        ("prior_width_vs", ctypes.c_double),
        ("prior_qs", ctypes.c_double),
        ("prior_width_qs", ctypes.c_double),
        ("prior_t_corr", ctypes.c_double),
        ("prior_width_t_corr", ctypes.c_double),
        ("prior_a_corr", ctypes.c_double),
        ("prior_width_a_corr", ctypes.c_double),
        ("step_size_z", ctypes.c_double),
        ("step_size_xy", ctypes.c_double),
        ("step_size_vs", ctypes.c_double),
        ("step_size_qs", ctypes.c_double),
        ("step_size_t_corr", ctypes.c_double),
        ("step_size_a_corr", ctypes.c_double),
        ("hist_xy_halfwidth", ctypes.c_double),
        ("hist_z_max", ctypes.c_double),
        ("abi_version", ctypes.c_int32),
        ("n_sta", ctypes.c_int32),
        ("n_events", ctypes.c_int32),
        ("n_procs", ctypes.c_int32),
        ("n_chains", ctypes.c_int32),
        ("n_cool", ctypes.c_int32),
        ("n_iter", ctypes.c_int32),
        ("n_burn", ctypes.c_int32),
        ("n_interval", ctypes.c_int32),
        ("solve_vs", ctypes.c_int32),
        ("solve_t_corr", ctypes.c_int32),
        ("solve_qs", ctypes.c_int32),
        ("solve_a_corr", ctypes.c_int32),
        ("use_time", ctypes.c_int32),
        ("use_amp", ctypes.c_int32),
        ("mode", ctypes.c_int32),
        ("precision", ctypes.c_int32),
        ("ladder", ctypes.c_int32),
        ("kernel", ctypes.c_int32),
        ("device", ctypes.c_int32),
        ("shard_rank", ctypes.c_int32),
        ("shard_count", ctypes.c_int32),
        ("hist_bins", ctypes.c_int32),
        ("max_samples", ctypes.c_int32),
        ("lane_slots", ctypes.c_int32),
        ("gibbs_shard_events", ctypes.c_int32),
        ("summary", ctypes.c_int32),
    ]


class StepTrace(ctypes.Structure):
    _fields_ = [
        ("proposal_type", ctypes.c_int32),
        ("index", ctypes.c_int32),
        ("prior_ok", ctypes.c_int32),
        ("accepted", ctypes.c_int32),
        ("log_likelihood", ctypes.c_double),
    ]


class SwapTrace(ctypes.Structure):
    _fields_ = [
        ("rank1", ctypes.c_int32),
        ("chain1", ctypes.c_int32),
        ("rank2", ctypes.c_int32),
        ("chain2", ctypes.c_int32),
        ("accepted", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


import numpy as _np

STEP_TRACE_DTYPE = _np.dtype(
    [("proposal_type", "<i4"), ("index", "<i4"), ("prior_ok", "<i4"), ("accepted", "<i4"),
     ("log_likelihood", "<f8")])
SWAP_TRACE_DTYPE = _np.dtype(
    [("rank1", "<i4"), ("chain1", "<i4"), ("rank2", "<i4"), ("chain2", "<i4"),
     ("accepted", "<i4"), ("reserved", "<i4")])


def default_config(**overrides):
    """sample/hypo_tremor.in:138-267 values; sizes must be supplied by the caller."""
    c = HtmConfig()
    c.abi_version = HTM_ABI_VERSION
    c.seed = 20231001
    c.temp_high = 200.0
    c.prior_z, c.prior_width_z, c.prior_width_xy = 0.0, 10.0, 30.0
    c.prior_vs, c.prior_width_vs = 3.0, 1.0
    c.prior_qs, c.prior_width_qs = 250.0, 100.0
    c.prior_t_corr, c.prior_width_t_corr = 0.0, 0.5
    c.prior_a_corr, c.prior_width_a_corr = 0.0, 0.02
    c.step_size_z, c.step_size_xy = 0.4, 2.0
    c.step_size_vs, c.step_size_qs = 0.2, 5.0
    c.step_size_t_corr, c.step_size_a_corr = 0.03, 0.005
    c.hist_xy_halfwidth, c.hist_z_max = 100.0, 60.0
    c.n_procs, c.n_chains, c.n_cool = 1, 5, 1
    c.n_iter, c.n_burn, c.n_interval = 4000000, 2000000, 1000
    c.solve_vs = c.solve_t_corr = c.solve_qs = c.solve_a_corr = 1
    c.use_time = c.use_amp = 1
    c.mode = MODE_BLOCKED_GIBBS
    c.precision = PRECISION_F32
    c.ladder = LADDER_RANDOM
    c.kernel = KERNEL_AUTO
    c.device = 0
    c.shard_rank, c.shard_count = 0, 1
    c.hist_bins = 0
    c.max_samples = 0
    for k, v in overrides.items():
        if not hasattr(c, k):
            raise AttributeError("htm_config has no field %r" % k)
        setattr(c, k, v)
    return c


def copy_config(cfg, **overrides):
    c = HtmConfig()
    ctypes.memmove(ctypes.byref(c), ctypes.byref(cfg), ctypes.sizeof(HtmConfig))
    for k, v in overrides.items():
        if not hasattr(c, k):
            raise AttributeError("htm_config has no field %r" % k)
        setattr(c, k, v)
    return c
