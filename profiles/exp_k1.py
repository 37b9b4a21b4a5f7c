"""Micro-experiments on the scoring / accumulation kernels (bench workload): kernel-only timings
under debug flags (option "debug_flags" is the debug word for the tcgen05 kernels)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from poccala_b200 import synth
from poccala_b200.engine import Corpus, Engine, EStep, Model

N_UNITS, MIX, N_UTT, T, L = 57, int(os.environ.get("MIX", 16)), int(os.environ.get("NUTT", 1000)), 300, 10
eng = Engine(0)
truth, init, labels, x = synth.torch_corpus(N_UTT, T, L, N_UNITS, MIX, 2, eng.device, 22)
corpus = Corpus(eng, labels, np.full(N_UTT, T, dtype=np.int32), N_UNITS)
model = Model(eng, *init, synth.default_transmat(N_UNITS))
es = EStep(eng, corpus, model)
es.load_frames(x)
es.score(); es.forward_backward(); es.accumulate()
torch.cuda.synchronize()
from poccala_b200 import _native as nat
from poccala_b200._native import _p
import ctypes as C
def t_kernel(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
def k1(): nat.call("pc_gmm_score", eng.h, corpus.c, _p(corpus.X), _p(model.W), MIX, _p(es.b), st())
def k3(): nat.call("pc_accumulate", eng.h, corpus.c, _p(corpus.X), _p(model.W), MIX, _p(es.b), _p(es.lgam), _p(es.acc), st())
for flags in [int(a) for a in sys.argv[1:]] or [0]:
    eng.set_option("debug_flags", flags)
    print("flags", flags, "K1 us %.1f" % (t_kernel(k1) if not (flags & 256) else 0.0), "K3 us %.1f" % t_kernel(k3), flush=True)
eng.set_option("debug_flags", 0)
print("K3 active (tile, unit) pairs: %d of %d" % (nat.lib().pc_corpus_active_tiles(corpus.c), nat.lib().pc_corpus_total_tiles(corpus.c)))
