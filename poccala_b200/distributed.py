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

On one NVLink / NVSwitch node the two collectives disappear altogether (PeerExchange): every rank's
statistics live in a block the other ranks map through CUDA IPC, and the M-step kernels read and add the
N copies themselves (csrc/peer.cu).  The process group is then only used to hand the IPC handles round.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np
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


class _DeviceDoubles:
    """A range of device memory owned by the library, for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 3,
                                         "strides": None}


class PeerExchange:
    """This rank's exchange block of the peer-memory reduction (pc_peer_*), connected to the blocks of all
    ranks of `group` (which must live on one node).  `views(which)` wraps a statistic set as torch tensors
    (acc [G,80], tsum [U,9], tmax [U,9]): which = 0 / 1 the sets of even / odd iterations, 2 the reduced sums
    of the last M-step."""

    def __init__(self, engine, n_units, n_gauss, group):
        from . import _native as nat

        self.engine, self.group = engine, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > nat.lib_const("PC_MAX_PEERS"):
            raise ValueError("at most %d ranks share an exchange" % nat.lib_const("PC_MAX_PEERS"))
        # Every step below is followed by the collective that comes next on ALL ranks, whatever happened locally: a rank
        # that cannot allocate, export or map a block must not leave the others waiting in a barrier.  The outcome is
        # agreed on at the end (MIN over the ranks) and a failure is raised everywhere.
        mine = np.zeros(64, dtype=np.uint8)
        err = None
        try:
            nat.call("pc_peer_create", engine.h, self.rank, self.world, int(n_gauss), int(n_units), nat._p(mine))
        except Exception as e:  # noqa: BLE001
            err = e
        # the handles travel through the process group (a uint8 tensor: NCCL and gloo both carry it)
        on_dev = dist.get_backend(group) == "nccl"
        t = torch.as_tensor(mine).to(engine.device if on_dev else "cpu")
        every = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(every, t, group=group)
        handles = np.ascontiguousarray(torch.stack(every).cpu().numpy())
        if err is None:
            try:
                nat.call("pc_peer_connect", engine.h, nat._p(handles))
            except Exception as e:  # noqa: BLE001
                err = e
        ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=engine.device if on_dev else "cpu")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)  # also the barrier: nobody signals into an unmapped block
        if int(ok.item()) == 0:
            try:
                nat.call("pc_peer_destroy", engine.h)
            except Exception:  # noqa: BLE001
                pass
            raise RuntimeError("peer-memory exchange unavailable (CUDA IPC between the ranks of this node): %s"
                               % (err if err is not None else "another rank failed"))
        self.n_units, self.n_gauss = int(n_units), int(n_gauss)
        self._views = [self._wrap(w) for w in range(3)]
        self._bound = weakref.WeakSet()  # EStep objects whose statistics live in this block (EStep.use_peer)

    def _wrap(self, which):
        from . import _native as nat

        ptrs = [C.c_void_p() for _ in range(3)]
        nat.call("pc_peer_buffers", self.engine.h, which, *[C.byref(p) for p in ptrs])
        sizes = (self.n_gauss * nat.KA, self.n_units * nat.TRANS_SLOTS, self.n_units * nat.TRANS_SLOTS)
        acc, tsum, tmax = (torch.as_tensor(_DeviceDoubles(p.value, n), device=self.engine.device)
                           for p, n in zip(ptrs, sizes))
        return acc.view(self.n_gauss, nat.KA), tsum.view(self.n_units, nat.TRANS_SLOTS), tmax.view(self.n_units, nat.TRANS_SLOTS)

    def views(self, which):
        return self._views[which]

    def current(self):
        """(acc, tsum, tmax) the next M-step will reduce."""
        return self._views[self.engine.get_option("peer_epoch") & 1]

    def timeouts(self):
        """Arrival waits that gave up since the last call (synchronises the device); 0 in a healthy run."""
        return self.engine.get_option("peer_timeouts")

    def close(self):
        from . import _native as nat

        for es in list(self._bound):  # nobody keeps a view of memory that is about to be unmapped
            es.use_nccl()
        self._views = None
        dist.barrier(group=self.group)  # every rank is past its last read
        nat.call("pc_peer_destroy", self.engine.h)
