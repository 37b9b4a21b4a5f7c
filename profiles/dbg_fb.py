"""Diagnostics for the scaled K2 kernel: which validity check sends utterances to the log kernel."""
import ctypes as C
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from poccala_b200 import synth
from poccala_b200.engine import Corpus, Engine, EStep, Model

eng = Engine(0)
def run(tag, n_utt, T, L, n_units, mix, seed, ragged):
    truth, init, labels, utts = synth.make_corpus(n_utt, T, L, n_units, mix, seed, ragged=ragged)
    tm = synth.default_transmat(n_units)
    corpus = Corpus(eng, labels, np.array([len(x) for x in utts], dtype=np.int32), n_units)
    model = Model(eng, *init, tm)
    es = EStep(eng, corpus, model)
    es.load_frames(torch.as_tensor(np.concatenate(utts)).to(eng.device))
    es.score(); es.forward_backward(); torch.cuda.synchronize()
    n = es.fb_fallbacks()
    flags = es.fb_flags()
    print(tag, "fallbacks", n, "of", n_utt, "reasons", dict(zip(*np.unique(flags, return_counts=True))))
    print("  T of flagged:", [len(utts[u]) for u in np.nonzero(flags)[0][:12]], "L:", [len(labels[u]) for u in np.nonzero(flags)[0][:12]])
    return es, corpus
run("test4", 20, 90, 5, 5, 4, 4, True)
run("cfg2-like", 200, 300, 10, 57, 16, 2, False)
