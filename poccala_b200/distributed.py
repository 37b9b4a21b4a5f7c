"""Multi-GPU part of one EM iteration: utterances shard across ranks, the only
exchange step is the reduction of the accumulators.  The reference merges timestamp-named .npy
accumulator files from every worker / machine with a log-sum-exp over the files (LHMM.py:256-290,
Clustering.py:314-367); here that is two collectives per iteration:

  1. all_reduce(MAX) of the per-unit maxima of the log-domain transition accumulators (Q6: values
     of magnitude 1e4..1e5, so the sum has to be formed relative to a global maximum),
  2. all_reduce(SUM) of ONE flat fp64 buffer holding the linear GMM statistics [G,80] and the
     transition sums exp(value - max) [U,9]  (a few MB: latency-bound, hence one call).

Every rank then runs the M-step on identical numbers: replicas stay bit-identical, no broadcast.
The functions take tensors on any device - NCCL over NVLink on the GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_utterances(n_utt, rank, world):
    """Utterance u -> rank u mod world (the reference deals utterances to Pool workers one at a
    time, AcousticModel.py:861-870)."""
    return list(range(rank, n_utt, world))


def allreduce_transition_maxima(tmax, group=None):
    """Collective 1: local maxima of the log-domain transition accumulators -> global maxima."""
    if group is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX, group=group)


def allreduce_flat_statistics(flat, group=None):
    """Collective 2: one SUM over [linear GMM statistics | transition sums relative to the global maxima]."""
    if group is not None:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)


def allreduce_em_statistics(flat, tmax, compute_tsum, group=None):
    """flat: 1-D fp64 tensor = [linear GMM statistics | transition sums]; the transition-sum part is
    (re)computed by `compute_tsum()` AFTER the maxima are global, then the whole buffer is summed.
    tmax: fp64 tensor of local maxima, replaced by the global ones.  No-op collectives when
    `group` is None (single rank).  (EStep issues the same three steps itself, the first two on a
    side stream so that they run under the accumulation kernel.)"""
    allreduce_transition_maxima(tmax, group)
    compute_tsum()
    allreduce_flat_statistics(flat, group)


def log_accumulators(tmax, tsum):
    """(max, sum) pair -> the reference's log-domain accumulator values."""
    return tmax + torch.log(tsum)
