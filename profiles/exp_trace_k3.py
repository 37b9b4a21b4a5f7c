import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
from poccala_b200 import synth, _native as nat
from poccala_b200.engine import Corpus, Engine, EStep, Model
N_UNITS, MIX, N_UTT, T, L = 57, 16, 1000, 300, 10
eng = Engine(0)
truth, init, labels, x = synth.torch_corpus(N_UTT, T, L, N_UNITS, MIX, 2, eng.device, 22)
corpus = Corpus(eng, labels, np.full(N_UTT, T, dtype=np.int32), N_UNITS)
model = Model(eng, *init, synth.default_transmat(N_UNITS))
es = EStep(eng, corpus, model); es.load_frames(x); es.score(); es.forward_backward(); torch.cuda.synchronize()
eng.set_option("debug_flags", 32)
for _ in range(2): es.accumulate()
torch.cuda.synchronize()
buf = (C.c_longlong * 8192)()
lib = nat.lib(); lib.pc_debug_read_acc.argtypes = [C.c_void_p, C.c_int]; lib.pc_debug_read_acc(buf, 8192)
blk = np.array(buf[4096:4096 + 148 * 4]).reshape(148, 4)
t0 = blk[:, 0].min()
print('per block: start us min/max %.1f %.1f | end us min/median/max %.1f %.1f %.1f | tiles min/median/max %d %d %d | items %d..%d' % ((blk[:,0].min()-t0)/1e3, (blk[:,0].max()-t0)/1e3, (blk[:,1].min()-t0)/1e3, np.median(blk[:,1]-t0)/1e3, (blk[:,1].max()-t0)/1e3, blk[:,2].min(), np.median(blk[:,2]), blk[:,2].max(), blk[:,3].min(), blk[:,3].max()))
print('ns per tile per block: min/median/max %.0f %.0f %.0f' % tuple(np.percentile((blk[:,1]-blk[:,0])/np.maximum(blk[:,2],1), [0,50,100])))
a = np.array(buf[:8000]).reshape(1000, 8)
n = 80
a = a[:n] - a[0, 0]
np.set_printoptions(linewidth=220)
print("cols: mma1_start, tile_landed, S_free | mma2_start, P_ready | smx_start, S_ready, P_written")
for i in range(0, 40): print(i, a[i])
d = np.diff(a[:, 0])
print("mean clk per tile (mma start to start): %.0f median %.0f" % (d.mean(), np.median(d)))
print("mma1: wait tile %.0f, wait S free %.0f, issue (to next start) %.0f | mma2: wait P %.0f, issue (to next start) %.0f" % ((a[:,1]-a[:,0]).mean(), (a[:,2]-a[:,1]).mean(), (a[1:,0]-a[:-1,2]).mean(), (a[:,4]-a[:,3]).mean(), (a[1:,3]-a[:-1,4]).mean()))
print("softmax: wait S %.0f, work %.0f" % ((a[:,6]-a[:,5]).mean(), (a[:,7]-a[:,6]).mean()))
