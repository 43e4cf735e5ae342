"""End-of-run gathers for event-sharded runs: one process per GPU, events
in contiguous blocks, no collective on the data path -- only these once-per-flush gathers of
posterior histograms, thinned samples and proposal counters.  NCCL over NVLink on GPUs; the same
code runs on gloo/CPU tensors in the tests.

The reference's only end-of-run reduction is the MPI_Reduce of the 7+7 counters
(src/cls_parallel.f90:265-268); histograms are new.
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from .synth import shard_bounds


class _CudaBuffer:
    """Expose a raw device pointer owned by libhtm_b200 through __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def histogram_tensor(handle):
    """Zero-copy torch view of the handle's device histograms: int32 [n_events(shard), 3, bins]
    (uint32 counts reinterpreted; they stay far below 2^31)."""
    ptr, nbytes = handle.device_ptr(0)
    shape = (handle.n_events, 3, handle.cfg.hist_bins)
    assert nbytes == int(np.prod(shape)) * 4
    return torch.as_tensor(_CudaBuffer(ptr, shape, "<i4"), device="cuda:%d" % handle.cfg.device)


def counts_tensor(handle):
    """Zero-copy torch view of the 14 device counters (int64: n_propose[7], n_accept[7])."""
    ptr, nbytes = handle.device_ptr(1)
    assert nbytes == 14 * 8
    return torch.as_tensor(_CudaBuffer(ptr, (14,), "<i8"), device="cuda:%d" % handle.cfg.device)


def gather_event_blocks(local, n_events_total, group=None):
    """all_gather of per-event blocks that were sharded with shard_bounds(): `local` is this rank's
    [n_events(shard), ...] tensor; returns the [n_events_total, ...] tensor on every rank.
    Shards may differ by one event, so blocks are padded to the largest shard for the collective."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_bounds(n_events_total, rank, world)
    if local.shape[0] != hi - lo:
        raise ValueError("rank %d holds %d events, shard_bounds says %d" % (rank, local.shape[0], hi - lo))
    biggest = max(shard_bounds(n_events_total, r, world)[1] - shard_bounds(n_events_total, r, world)[0]
                  for r in range(world))
    tail = tuple(local.shape[1:])
    padded = torch.zeros((biggest,) + tail, dtype=local.dtype, device=local.device)
    padded[: hi - lo] = local
    out = torch.empty((world, biggest) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(world * biggest, *tail), padded, group=group)
    parts = []
    for r in range(world):
        a, b = shard_bounds(n_events_total, r, world)
        parts.append(out[r, : b - a])
    return torch.cat(parts, dim=0)


def reduce_counts(local_counts, group=None):
    """Sum of the proposal / acceptance counters over shards (MPI_Reduce of cls_parallel.f90:265-268,
    as an all-reduce so every rank can write proposal_count.txt)."""
    total = local_counts.clone()
    dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return total


def gather_run(handle, group=None):
    """Histograms of every event and summed counters, on every rank, from a sharded GPU run."""
    handle.synchronize()
    hist = gather_event_blocks(histogram_tensor(handle), handle.cfg.n_events, group)
    counts = reduce_counts(counts_tensor(handle), group)
    return hist, counts
