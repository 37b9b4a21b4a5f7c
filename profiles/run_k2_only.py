"""cfg2 corpus, K2 only (for ncu): python profiles/run_k2_only.py [n_utt]"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from poccala_b200 import synth
from poccala_b200.engine import Corpus, Engine, EStep, Model
n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
eng = Engine(0)
truth, init0, labels, x = synth.torch_corpus(n_utt, 300, 10, 57, 16, 2, eng.device, 22)
corpus = Corpus(eng, labels, np.full(n_utt, 300, dtype=np.int32), 57)
model = Model(eng, *init0, synth.default_transmat(57))
es = EStep(eng, corpus, model)
es.load_frames(x)
es.score()
for _ in range(5):
    es.forward_backward()
torch.cuda.synchronize()
