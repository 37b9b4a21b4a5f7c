"""K1 timing with CUDA events: python profiles/time_k1.py [n_utt] [mix] ; prints ms per launch for k1_kernel 1 and 0"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from poccala_b200 import synth
from poccala_b200.engine import Corpus, Engine, EStep, Model
n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
mix = int(sys.argv[2]) if len(sys.argv) > 2 else 16
eng = Engine(0)
truth, init0, labels, x = synth.torch_corpus(n_utt, 300, 10, 57, mix, 2, eng.device, 22)
corpus = Corpus(eng, labels, np.full(n_utt, 300, dtype=np.int32), 57)
model = Model(eng, *init0, synth.default_transmat(57))
es = EStep(eng, corpus, model)
es.load_frames(x)
ref = None
import os
for k in (1, 0):
    eng.set_option("k1_kernel", k)
    eng.set_option("debug_flags", int(os.environ.get("K1_DBG", "0")))
    for _ in range(2):
        es.score()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        es.score()
    e1.record()
    torch.cuda.synchronize()
    out = es.b.clone()
    if ref is None:
        ref = out
    print("k1_kernel", k, "mix", mix, "utt", n_utt, "ms", e0.elapsed_time(e1) / 5, "max diff vs k1_kernel=1", float((out - ref).abs().max()), flush=True)
