import random, sys
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import fast
from poccala_b200 import synth
from poccala_b200.engine import Corpus, Engine, EStep, Model, kmeans_run, kmeans_seed_points
from tests.test_gpu_configs import _uniform_state_data
eng = Engine(0)
n_units, mix = 4, 8
truth, _, labels, utts = synth.make_corpus(48, 150, 4, n_units, mix, 5)
data = _uniform_state_data(labels, utts, n_units)
flat, seeds, off = [], [], [0]
for u in range(n_units):
    for r in range(3):
        x = data[u][r]; rnd = random.Random(100 * u + r)
        seeds.append(kmeans_seed_points(np.ascontiguousarray(x[:, 0]), mix, rnd)); flat.append(x); off.append(off[-1] + len(x))
out = kmeans_run(eng, torch.as_tensor(np.concatenate(flat)).to(eng.device), np.array(off), mix, np.array(seeds, dtype=np.int32))
mean = out["mean"].cpu().numpy().reshape(n_units, 3, mix, 39); var = out["var"].cpu().numpy().reshape(n_units, 3, mix, 39)
alpha = out["alpha"].cpu().numpy().reshape(n_units, 3, mix); alpha = alpha / alpha.sum(-1, keepdims=True)
init = (mean, np.maximum(var, 1e-4), alpha)
tm = synth.default_transmat(n_units)
om = fast.Model(*init, tm)
gvar = np.concatenate(utts, axis=0).var(axis=0)
for tc in (1, 0):
    eng.set_option("tensor_core", tc)
    corpus = Corpus(eng, labels, np.array([len(x) for x in utts], dtype=np.int32), n_units)
    model = Model(eng, *init, tm); es = EStep(eng, corpus, model)
    es.load_frames(torch.as_tensor(np.concatenate(utts)).to(eng.device))
    o2 = fast.Model(*init, tm)
    for it in range(2):
        es.em_iteration(c_covariance=1e-6); torch.cuda.synchronize()
        stats, info = fast.estep_corpus(o2, labels, utts); new = fast.mstep(o2, stats, c_covariance=1e-6)
        m, v, a, t = model.numpy()
        ok = stats.occ >= 1e-4
        rv = (np.abs(v - new.var) / (1e-4 * np.maximum(new.var, 1e-2 * gvar)))
        rm = (np.abs(m - new.mean) / (1e-4 * np.maximum(np.abs(new.mean), np.sqrt(new.var))))
        rv[~ok] = 0; rm[~ok] = 0
        i = np.unravel_index(np.argmax(rv), rv.shape)
        print("tc", tc, "iter", it, "max var ratio %.3f at %s occ %.3e var %.3e ref %.3e ; max mean ratio %.3f" % (rv.max(), i, stats.occ[i[:3]], v[i], new.var[i], rm.max()))
        o2 = new
