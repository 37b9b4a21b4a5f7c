"""One corpus, a few E-steps (for ncu): python profiles/run_estep_only.py [n_utt] [mix]"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from poccala_b200 import synth
from poccala_b200.engine import Corpus, Engine, EStep, Model
n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
mix = int(sys.argv[2]) if len(sys.argv) > 2 else 64
eng = Engine(0)
truth, init0, labels, x = synth.torch_corpus(n_utt, 300, 10, 57, mix, 2, eng.device, 22)
corpus = Corpus(eng, labels, np.full(n_utt, 300, dtype=np.int32), 57)
model = Model(eng, *init0, synth.default_transmat(57))
es = EStep(eng, corpus, model)
es.load_frames(x)
for _ in range(4):
    es.estep()
torch.cuda.synchronize()
