"""Multi-rank parity on hardware: the same corpus trained on one rank and sharded over N NCCL ranks must
give the same model, and the replicas must stay bit-identical (AcousticModel.py:842-882 merges the
accumulators of all workers).  The N-rank run needs N GPUs on the box
(`gpurun --gpus 2`); with one GPU the single-rank variant still checks the host-buffer entry point against
the device-resident path."""
import json
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_check(n):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    if n == 1:
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "1", "--check"]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
               "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "bench.py"), "--gpus", str(n), "--check"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    return json.loads(lines[-1])


def test_single_rank_check_mode():
    out = _run_check(1)
    assert out["check"] == "ok", out
    assert max(out["host_entry_diff_vs_single_rank_iteration_1"].values()) <= 1e-9, out  # same data, same order


@pytest.mark.parametrize("n", [2, 4, 8])
def test_sharded_training_matches_single_rank_nccl(n):
    if torch.cuda.device_count() < n:
        pytest.skip("needs %d GPUs" % n)
    out = _run_check(n)
    assert out["check"] == "ok", out
    assert out["replicas_bit_identical"], out
    # the shards regroup the fp32 partial sums of the accumulation kernel: 5e-5 after one iteration from the same model
    # (measured on a variance 1.6e-5 at 2 ranks, 3.0e-5 at 8; 7e-7 on the means), for the device-resident path, the host
    # entry point with the NCCL hook and the host entry point over peer memory alike
    assert max(out["diff_vs_single_rank_iteration_1"].values()) <= 5e-5, out
    assert max(out["host_entry_diff_vs_single_rank_iteration_1"].values()) <= 5e-5, out
    assert max(out["peer_memory_host_entry_diff_vs_single_rank_iteration_1"].values()) <= 5e-5, out
    assert max(out["diff_vs_single_rank_iteration_2"].values()) <= 1e-3, out
    # the peer-memory reduction gives the NCCL reduction's model (sums of N terms in another order, amplified over six
    # iterations: measured 4e-10 at 2 ranks), no waits gave up
    assert max(out["peer_memory_vs_nccl_iteration_6"].values()) <= 1e-7 and out["peer_timeouts"] == 0, out
