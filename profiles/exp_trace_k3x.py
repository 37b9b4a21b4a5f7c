"""Block 0's timeline in the gathered-block accumulation kernel (debug_flags & 32): python profiles/exp_trace_k3x.py [mix]"""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
from poccala_b200 import synth, _native as nat
from poccala_b200.engine import Corpus, Engine, EStep, Model
mix = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n_utt, T, L = 1000, 300, 10
eng = Engine(0)
truth, init0, labels, x = synth.torch_corpus(n_utt, T, L, 57, mix, 2, eng.device, 22)
corpus = Corpus(eng, labels, np.full(n_utt, T, dtype=np.int32), 57)
model = Model(eng, *init0, synth.default_transmat(57))
es = EStep(eng, corpus, model)
es.load_frames(x)
es.score()
for _ in range(2):
    es.forward_backward(); es.acc.zero_(); es.accumulate()
es.forward_backward(); es.acc.zero_()
eng.set_option("debug_flags", 32)
es.accumulate()
torch.cuda.synchronize()
eng.set_option("debug_flags", 0)
buf = (C.c_longlong * 8192)()
lib = nat.lib(); lib.pc_debug_read_k3x.argtypes = [C.c_void_p, C.c_int]; lib.pc_debug_read_k3x(buf, 8192)
a = np.array(list(buf), dtype=np.int64)
mma = a[:2400].reshape(-1, 8); sm = a[2400:4800].reshape(-1, 8); pr = a[4800:7200].reshape(-1, 8)
nb = int((mma[:, 3] > 0).sum())
t0 = mma[0, 0]
print("batches of block 0:", nb, "clk per batch:", (mma[nb - 1, 4] - t0) / max(nb, 1))
print("  n | MMA: start(rel) wait-img wait-S issue1 issue2 | softmax: wait-S read wait-P write | prepare: start(rel) loads wait-img wait-stg convert")
for n in range(min(nb, 40)):
    m, s_, p = mma[n], sm[n], pr[n]
    print("%3d | %8d %6d %6d %6d %6d | %6d %5d %6d %6d | %8d %6d %6d %6d %6d" % (
        n, m[0] - t0, m[1] - m[0], m[2] - m[1], m[3] - m[2], m[4] - m[3],
        s_[1] - s_[0], s_[2] - s_[1], s_[3] - s_[2], s_[4] - s_[3],
        p[0] - t0, p[1] - p[0], p[2] - p[1], p[3] - p[2], p[4] - p[3]))
