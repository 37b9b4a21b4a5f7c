"""N > 1 path on CPU (gloo, world size 2): utterance sharding + the two-collective accumulator
reduction of poccala_b200.distributed give every rank the single-rank statistics, hence the same
re-estimated model.  The per-shard numbers come from the CPU oracle; on the GPUs the
same function runs on the kernels' buffers over NCCL (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from oracle import fast  # noqa: E402
from poccala_b200 import synth  # noqa: E402
from poccala_b200.distributed import allreduce_em_statistics, log_accumulators, shard_utterances  # noqa: E402

N_UNITS, MIX, N_UTT = 4, 4, 10


def _corpus():
    truth, init, labels, utts = synth.make_corpus(N_UTT, 40, 3, N_UNITS, MIX, 61, ragged=True)
    return init, labels, utts


def _pack(stats):
    """oracle statistics -> (flat [linear | transition sums], tmax, log values) as the engine lays them out."""
    stats.finalize()
    lin = np.concatenate([stats.sx, stats.occ[..., None], stats.sxx, stats.occ[..., None]], axis=-1)  # [U,3,M,80]
    logv = np.full((N_UNITS, 9), -np.inf)
    for r in range(3):
        logv[:, 3 * r + 0] = stats.ksai_acc[:, r, r + 1]
        logv[:, 3 * r + 1] = stats.ksai_acc[:, r, r + 2]
        logv[:, 3 * r + 2] = stats.gamma_acc[:, r]
    return lin.reshape(-1), logv


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    init, labels, utts = _corpus()
    om = fast.Model(*init, synth.default_transmat(N_UNITS))
    mine = shard_utterances(N_UTT, rank, world)
    stats, _ = fast.estep_corpus(om, [labels[u] for u in mine], [utts[u] for u in mine])
    lin, logv = _pack(stats)
    n_lin = lin.size
    flat = torch.zeros(n_lin + logv.size, dtype=torch.float64)
    flat[:n_lin] = torch.as_tensor(lin)
    tmax = torch.as_tensor(logv.copy())

    def local_sums():
        with np.errstate(invalid="ignore"):
            t = np.exp(logv - tmax.numpy())
        flat[n_lin:] = torch.as_tensor(np.where(np.isfinite(logv), t, 0.0).reshape(-1))

    allreduce_em_statistics(flat, tmax, local_sums, dist.group.WORLD)
    logsum = log_accumulators(tmax, flat[n_lin:].view(N_UNITS, 9))
    out[rank] = (flat[:n_lin].numpy().copy(), logsum.numpy().copy())
    dist.destroy_process_group()


def test_two_rank_reduction_equals_single_rank():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    init, labels, utts = _corpus()
    om = fast.Model(*init, synth.default_transmat(N_UNITS))
    stats, _ = fast.estep_corpus(om, labels, utts)
    lin_ref, log_ref = _pack(stats)
    assert sorted(shard_utterances(N_UTT, 0, 2) + shard_utterances(N_UTT, 1, 2)) == list(range(N_UTT))
    for rank in (0, 1):
        lin, logsum = out[rank]
        assert np.allclose(lin, lin_ref, rtol=1e-12, atol=1e-12)
        fin = np.isfinite(log_ref)
        assert (np.isfinite(logsum) == fin).all()
        assert np.abs(logsum[fin] - log_ref[fin]).max() < 1e-9
    # both ranks hold identical numbers -> identical M-step, no broadcast needed
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
