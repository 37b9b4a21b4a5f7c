import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
from poccala_b200 import synth, _native as nat
from poccala_b200._native import _p
from poccala_b200.engine import Corpus, Engine, EStep, Model
N_UNITS, MIX, N_UTT, T, L = 57, 16, 1000, 300, 10
eng = Engine(0)
truth, init, labels, x = synth.torch_corpus(N_UTT, T, L, N_UNITS, MIX, 2, eng.device, 22)
corpus = Corpus(eng, labels, np.full(N_UTT, T, dtype=np.int32), N_UNITS)
model = Model(eng, *init, synth.default_transmat(N_UNITS))
es = EStep(eng, corpus, model); es.load_frames(x); es.score(); torch.cuda.synchronize()
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 32
eng.set_option("debug_flags", flags)
for _ in range(3):
    nat.call("pc_gmm_score", eng.h, corpus.c, _p(corpus.X), _p(model.W), MIX, _p(es.b), C.c_void_p(0))
torch.cuda.synchronize()
buf = (C.c_longlong * 8000)()
lib = nat.lib(); lib.pc_debug_read.argtypes = [C.c_void_p, C.c_int]; lib.pc_debug_read(buf, 8000)
a = np.array(buf[:]).reshape(1000, 8)
npairs = 210
t0 = a[0, 0]
a = a[:npairs] - t0
np.set_printoptions(linewidth=200)
print("cols: mma_start, after_a_full, after_tm_empty, after_issue | epi_start, epi_woke, epi_done")
for i in list(range(0, 36)) + list(range(90, 100)):
    print(i, a[i, :7])
d = np.diff(a[:, 0])
print("mean cycles per pair (mma start to start):", d.mean(), "median", np.median(d))
print("mma: wait a_full", (a[:,1]-a[:,0]).mean(), "wait tm_empty", (a[:,2]-a[:,1]).mean(), "issue", (a[:,3]-a[:,2]).mean())
print("epi: wait", (a[:,5]-a[:,4]).mean(), "work", (a[:,6]-a[:,5]).mean(), "issue->woke", (a[:,5]-a[:,3]).mean())
