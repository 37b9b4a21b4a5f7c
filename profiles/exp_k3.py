"""K3 timing on the bench workload, both tensor-core kernels: python profiles/exp_k3.py [n_utt] [mix]"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from poccala_b200 import synth
from poccala_b200.engine import Corpus, Engine, EStep, Model
n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
mix = int(sys.argv[2]) if len(sys.argv) > 2 else 16
T, L = 300, 10
eng = Engine(0)
truth, init0, labels, x = synth.torch_corpus(n_utt, T, L, 57, mix, 2, eng.device, 22)
corpus = Corpus(eng, labels, np.full(n_utt, T, dtype=np.int32), 57)
model = Model(eng, *init0, synth.default_transmat(57))
es = EStep(eng, corpus, model)
es.load_frames(x)
es.score()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.device)
ref = None
for name, k3, dbg in (("per tile", 0, 0), ("gathered blocks", 1, 0), ("gathered, no MMA", 1, 2), ("gathered, P = 0", 1, 16),
                      ("gathered, no conversion", 1, 8)):
    eng.set_option("k3_kernel", k3)
    eng.set_option("debug_flags", dbg)
    ms = []
    for it in range(6):
        es.forward_backward()   # sets the activity masks for the accumulation that follows
        es.acc.zero_()
        flush.zero_()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record(); es.accumulate(); ev[1].record()
        torch.cuda.synchronize()
        if it:
            ms.append(ev[0].elapsed_time(ev[1]))
    if dbg == 0:
        a = es.acc.clone()
        if ref is None:
            ref = a
        else:
            d = (a - ref).abs()
            print("   max |acc - acc_ref| = %.3g (rel. to max |acc| %.3g), occupancy sums %.6f vs %.6f" % (
                d.max().item(), d.max().item() / ref.abs().max().item(), a[:, 39].sum().item(), ref[:, 39].sum().item()))
    pairs = n_utt * T * 3 * L * mix
    print("%-28s %.1f us (min %.1f)  %.0f TFLOP/s algorithmic" % (name, 1e3 * sum(ms) / len(ms), 1e3 * min(ms),
                                                                   158.0 * pairs / (min(ms) * 1e-3) / 1e12), flush=True)
eng.set_option("debug_flags", 0)
