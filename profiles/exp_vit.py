"""K4 timing and block 0 phase clocks at the configs[3] shape: python profiles/exp_vit.py [n_utt]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poccala_b200 import _native as nat, synth  # noqa: E402
from poccala_b200.engine import Corpus, Engine, host_log_bands, viterbi  # noqa: E402

U = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
T, L, NU = 1000, 20, 57
eng = Engine(0)
rng = np.random.default_rng(0)
labels = rng.integers(0, NU, size=(U, L)).astype(np.int32)
corpus = Corpus(eng, labels, np.full(U, T, dtype=np.int32), NU)
b = (-60.0 - 20.0 * torch.rand(corpus.emis_floats, device="cuda")).float()
ls, ln = host_log_bands(synth.default_transmat(NU), eng.device)
N = 3 * L + 2
logpi = torch.full((U,), float(np.log(1.0 / N)), dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(2):
    viterbi(eng, corpus, b, ls, ln, utt_logpi=logpi)
ts = []
for _ in range(5):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    viterbi(eng, corpus, b, ls, ln, utt_logpi=logpi)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
buf = (C.c_longlong * 8)()
nat.lib().pc_debug_read_vit.argtypes = [C.c_void_p]
nat.lib().pc_debug_read_vit(buf)
print("K4: %.1f us (median of 5), %d utt x %d frames; block 0: recurrence %d clk (%.0f per frame), traceback %d clk (%.0f per frame)"
      % (np.median(ts) * 1e3, U, T, buf[1] - buf[0], (buf[1] - buf[0]) / T, buf[2] - buf[1], (buf[2] - buf[1]) / T))
