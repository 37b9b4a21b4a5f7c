import sys
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import fast
from poccala_b200 import synth
from poccala_b200.engine import Corpus, Engine, EStep, Model
from tests.helpers import load_golden
np.set_printoptions(precision=6, linewidth=200)
eng = Engine(0)
g = load_golden("estep_small.npz")
n = int(g["n_utt"])
labels = [g[f"u{k}_label"] for k in range(n)]
utts = [g[f"u{k}_X"] for k in range(n)]
tm0 = synth.default_transmat(3)
om = fast.Model(g["mean"], g["var"], g["alpha"], tm0)
res = {}
for exact in (1, 0):
    corpus = Corpus(eng, labels, np.array([len(x) for x in utts], dtype=np.int32), 3)
    model = Model(eng, g["mean"], g["var"], g["alpha"], tm0)
    es = EStep(eng, corpus, model)
    es.load_frames(torch.as_tensor(np.concatenate(utts)).to(eng.device))
    eng.set_option("fb_exact", exact)
    es.score(); es.forward_backward(); torch.cuda.synchronize()
    print("exact", exact, "fallbacks", es.fb_fallbacks(), "logp", es.utt_logp.cpu().numpy(), "iters", es.utt_iters.cpu().numpy())
    res[exact] = (es.pair_trans.cpu().numpy().copy(), es.utt_logp.cpu().numpy().copy(), es.lgam.cpu().numpy().copy())
    es.accumulate(); es.reduce_transitions(); es.mstep(c_covariance=1e-6); torch.cuda.synchronize()
    tmn = model.numpy()[3]
    print(" max |tm - golden|", np.abs(tmn - g["it1_transmat"]).max(), "\n", tmn[1], "\n", tmn[2])
print("golden\n", g["it1_transmat"][1], "\n", g["it1_transmat"][2])
pe, pf = res[1][0], res[0][0]
print("pair_trans exact\n", pe[:6]); print("pair_trans scaled\n", pf[:6])
le, lf = res[1][2], res[0][2]
d = np.abs(np.exp(le.astype(np.float64)) - np.exp(lf.astype(np.float64)))
print("max gamma diff", np.nanmax(d), "at", np.nanargmax(d), "T", [len(x) for x in utts], "L", [len(l) for l in labels])
