"""K2 timing and block 0 phase clocks on the bench workload: python profiles/exp_k2.py"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from poccala_b200 import synth
from poccala_b200.engine import Corpus, Engine, EStep, Model
N_UNITS, N_INITIALS, MIX, N_UTT, T, L = 57, 22, 16, 1000, 300, 10
eng = Engine(0)
truth, init0, labels, x = synth.torch_corpus(N_UTT, T, L, N_UNITS, MIX, 2, eng.device, N_INITIALS)
corpus = Corpus(eng, labels, np.full(N_UTT, T, dtype=np.int32), N_UNITS)
model = Model(eng, *init0, synth.default_transmat(N_UNITS))
es = EStep(eng, corpus, model)
es.load_frames(x)
es.score()
ref = None
for cfg in [0]:
    for _ in range(3):
        es.forward_backward()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ls, ln = model.log_bands()
    from poccala_b200 import _native as nat
    from poccala_b200.engine import _p, _stream
    ev[0].record()
    for _ in range(10):
        nat.call("pc_forward_backward", eng.h, corpus.c, _p(es.b), _p(ls), _p(ln), _p(es.lgam), _p(es.utt_logp),
                 _p(es.utt_iters), _p(es.pair_trans), _stream())
    ev[1].record()
    torch.cuda.synchronize()
    lp = es.utt_logp.sum().item()
    print("K2: %.1f us per call, sum logp %.6f" % (ev[0].elapsed_time(ev[1]) * 100, lp), flush=True)

import ctypes as C
buf = (C.c_longlong * 16)()
lib = nat.lib(); lib.pc_debug_read_fb.argtypes = [C.c_void_p]; lib.pc_debug_read_fb(buf)
a = list(buf)
print("block 0 clocks: backward %d, pi %d, forward %d, helper tail after forward %d (T=%d: %.0f / %.0f clk per frame)" % (
    a[1] - a[0], a[2] - a[1], a[3] - a[2], a[9] - a[3], T, (a[1] - a[0]) / (T - 1), (a[3] - a[2]) / (T - 1)))
