"""Block 0's timeline in the wide scoring kernel (debug_flags & 32): python profiles/exp_trace_k1w.py"""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
from poccala_b200 import synth, _native as nat
from poccala_b200.engine import Corpus, Engine, EStep, Model
n_utt, T, L, mix = 1000, 300, 10, 16
eng = Engine(0)
truth, init0, labels, x = synth.torch_corpus(n_utt, T, L, 57, mix, 2, eng.device, 22)
corpus = Corpus(eng, labels, np.full(n_utt, T, dtype=np.int32), 57)
model = Model(eng, *init0, synth.default_transmat(57))
es = EStep(eng, corpus, model)
es.load_frames(x)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.device)
for _ in range(3):
    es.score()
eng.set_option("debug_flags", 32)
flush.zero_()
es.score()
torch.cuda.synchronize()
eng.set_option("debug_flags", 0)
buf = (C.c_longlong * 8192)()
lib = nat.lib(); lib.pc_debug_read_k1w.argtypes = [C.c_void_p, C.c_int]; lib.pc_debug_read_k1w(buf, 8192)
a = np.array(list(buf), dtype=np.int64)
acc = a[:4096].reshape(-1, 4); conv = a[4096:6144].reshape(-1, 4); mw = a[6144:6656]; epi = a[6656:6656 + 1500].reshape(-1, 5)
t0 = acc[0, 0]
n_acc = int((acc[:, 3] > 0).sum())
print("accumulators recorded:", n_acc, " total clk:", acc[n_acc - 1, 3] - t0)
print(" acc | start(rel) | wait B | wait TMEM | issue | gap to next start")
for n in range(min(n_acc, 42)):
    s, b_, t_, i_ = acc[n]
    nxt = acc[n + 1, 0] - i_ if n + 1 < n_acc else 0
    print("%4d | %9d | %6d | %8d | %5d | %6d" % (n, s - t0, b_ - s, t_ - b_, i_ - t_, nxt))
print(" tile | conv start(rel) | wait rows | wait operand buffer | convert | MMA-warp wait for image (start rel)")
for n in range(min(int((conv[:, 3] > 0).sum()), 21)):
    c0, c1, c2, c3 = conv[n]
    print("%4d | %9d | %6d | %8d | %5d | %9d" % (n, c0 - t0, c1 - c0, c2 - c1, c3 - c2, mw[n] - t0))
print(" acc | epilogue warp 0: wait start(rel) | wait tm_full | convert | math | barrier | (next wait start - barrier end = copy-out)")
for n in range(min(n_acc, 24)):
    e0, e1, e2, e3, e4 = epi[n]
    print("%4d | %9d | %6d | %6d | %6d | %6d | %6d" % (n, e0 - t0, e1 - e0, e2 - e1, e3 - e2, e4 - e3, epi[n + 1, 0] - e4))
