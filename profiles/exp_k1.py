"""K1 timing on the bench workload, both tensor-core kernels: python profiles/exp_k1.py [n_utt] [mix]"""
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.device)
ref = None
for name, k1, dbg in (("per-position-pair", 0, 0), ("wide", 1, 0), ("wide, no MMA", 1, 2), ("wide, no epilogue math", 1, 16),
                      ("wide, no conversion", 1, 8), ("wide, no MMA/epilogue/conversion", 1, 26)):
    eng.set_option("k1_kernel", k1)
    eng.set_option("debug_flags", dbg)
    for _ in range(3):
        es.score()
    torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        flush.zero_()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record(); es.score(); ev[1].record()
        torch.cuda.synchronize()
        ms.append(ev[0].elapsed_time(ev[1]))
    if dbg == 0:
        bb = es.b.clone()
        if ref is None:
            ref = bb
        else:
            print("   max |b - b_ref| = %.3g" % (bb - ref).abs().max().item())
    pairs = n_utt * T * 3 * L * mix
    print("%-36s %.1f us (min %.1f)  %.0f TFLOP/s algorithmic" % (name, 1e3 * sum(ms) / len(ms), 1e3 * min(ms),
                                                                   158.0 * pairs / (min(ms) * 1e-3) / 1e12), flush=True)
eng.set_option("debug_flags", 0)
