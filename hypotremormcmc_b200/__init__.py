"""B200-native (sm_100a) implementation of the hypo_tremor_mcmc inversion hot path of
akuhara/HypoTremorMCMC: cls_forward likelihood, cls_mcmc perturb/accept, cls_parallel
tempering swap, behind the C ABI of include/htm_b200.h.

This package is the Python host-side mirror used by tests and bench.py; the product is
``csrc/libhtm_b200.so``.  There is no CPU fallback: importing works without a GPU (so the
symbol table can be checked), every compute call needs one.
"""
from .config import (HtmConfig, StepTrace, SwapTrace, default_config, copy_config,  # noqa: F401
                     MODE_REPLAY, MODE_FACTORISED, MODE_BLOCKED_GIBBS, PRECISION_F64,
                     PRECISION_F32, LADDER_RANDOM, LADDER_GEOMETRIC, KERNEL_AUTO,
                     KERNEL_WARP_PER_CHAIN, KERNEL_LANE_PER_CHAIN, PROPOSAL_LABELS,
                     STEP_TRACE_DTYPE, SWAP_TRACE_DTYPE)
from .synth import Synthetic, shard_bounds  # noqa: F401
from . import io  # noqa: F401
from .api import HtmError, HypoTremorB200, load_library, library_path  # noqa: F401

__version__ = "0.1.0"
