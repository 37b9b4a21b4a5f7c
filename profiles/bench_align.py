"""Alignment post-processing at the configs[4] per-GPU shard shape (R = 8: 12 500
utterances x 300 frames, 10 units of 57 each, 171 data sets): device time of the three steps and their
HBM roofline fractions.  Usage: python profiles/bench_align.py [n_utt]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poccala_b200.engine import Corpus, Engine, _p, _stream, group_frames, segment_keys  # noqa: E402
from poccala_b200 import _native as nat  # noqa: E402

U = int(sys.argv[1]) if len(sys.argv) > 1 else 12500
T, L, NU = 300, 10, 57
rng = np.random.default_rng(0)
eng = Engine(0)
labels = rng.integers(0, NU, size=(U, L)).astype(np.int32)
corpus = Corpus(eng, labels, np.full(U, T, dtype=np.int32), NU)
F = U * T
# a monotone alignment per utterance: 3L cut points
cuts = np.sort(rng.integers(1, T, size=(U, 3 * L - 1)), axis=1)
path = (1 + (np.arange(T)[None, :, None] >= cuts[:, None, :]).sum(-1)).astype(np.int32).reshape(-1)
path_d = torch.as_tensor(path).cuda()
x = torch.randn(F, 39, dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=5):
    ts = []
    for _ in range(reps + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts[2:]))


peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {}
hbm = float(peaks.get("hbm_gbs", 6541.5))
key = torch.empty(F, dtype=torch.int32, device="cuda")
kept = torch.empty(U, dtype=torch.int32, device="cuda")
out = {"frames": F, "keys": NU * 3, "hbm_peak_gbps": hbm}
for mode, name in ((0, "segment_uniform"), (1, "segment_aligned")):
    ms = timed(lambda: nat.call("pc_segment_keys", eng.h, corpus.c, mode, _p(path_d) if mode else None, _p(key), _p(kept), _stream()))
    byts = F * (8 if mode else 4)
    out[name] = {"ms": ms, "GBps": byts / ms / 1e6, "frac": byts / ms / 1e6 / hbm}
ws = torch.empty(int(nat.lib().pc_group_workspace_bytes(F, NU * 3)), dtype=torch.uint8, device="cuda")
key_off = torch.empty(NU * 3 + 2, dtype=torch.int64, device="cuda")
order = torch.empty(F, dtype=torch.int32, device="cuda")
ms = timed(lambda: nat.call("pc_group_frames", eng.h, _p(key), F, NU * 3, _p(ws), _p(key_off), _p(order), _stream()))
out["group"] = {"ms": ms, "GBps": 12 * F / ms / 1e6, "frac": 12 * F / ms / 1e6 / hbm}
dst = torch.empty_like(x)
n_kept = int(key_off[NU * 3])
ms = timed(lambda: nat.call("pc_gather_rows", eng.h, _p(order), n_kept, 39 * 8, _p(x), _p(dst), _stream()))
out["gather"] = {"ms": ms, "rows": n_kept, "GBps": 2 * 312 * n_kept / ms / 1e6, "frac": 2 * 312 * n_kept / ms / 1e6 / hbm}
print(json.dumps(out))
