"""Block 0's timeline in the 64-mixture scoring kernel (debug_flags & 32): python profiles/exp_trace_k1b.py [n_utt]"""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
from poccala_b200 import synth, _native as nat
from poccala_b200.engine import Corpus, Engine, EStep, Model
n_utt, T, L, mix = int(sys.argv[1]) if len(sys.argv) > 1 else 4000, 300, 10, 64
eng = Engine(0)
truth, init0, labels, x = synth.torch_corpus(n_utt, T, L, 57, mix, 2, eng.device, 22)
corpus = Corpus(eng, labels, np.full(n_utt, T, dtype=np.int32), 57)
model = Model(eng, *init0, synth.default_transmat(57))
es = EStep(eng, corpus, model)
es.load_frames(x)
for _ in range(3):
    es.score()
eng.set_option("debug_flags", 32)
es.score()
torch.cuda.synchronize()
eng.set_option("debug_flags", 0)
buf = (C.c_longlong * 8192)()
lib = nat.lib(); lib.pc_debug_read_k1b.argtypes = [C.c_void_p, C.c_int]; lib.pc_debug_read_k1b(buf, 8192)
a = np.array(list(buf), dtype=np.int64)
m = a[:4800].reshape(-1, 8); e = a[4800:4800 + 2400].reshape(-1, 4)
t0 = m[0, 0]
print(" acc | start(rel) | hi wait | tile wait | TMEM wait | lo wait (first) | lo wait (late) | issue | epi: wait full | math | epi start rel")
for n in range(3, 3 + 66):
    s, a1, a2, a3, l0, l1, iss, hw = m[n]
    hiw = (s - hw) if hw > 0 else 0
    late = (l1 - l0) if l0 > 0 else 0
    print("%4d | %9d | %6d | %6d | %6d | %6d | %6d | %6d | %6d | %6d | %9d" % (n, s - t0, hiw, a1 - s, a2 - a1, a3 - a2, late, iss - a3 - late, e[n, 1] - e[n, 0], e[n, 2] - e[n, 1], e[n, 0] - t0))
n_rec = 590
print("clk per accumulator over %d:" % n_rec, (m[n_rec, 0] - m[30, 0]) / (n_rec - 30))
