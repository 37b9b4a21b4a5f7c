"""Achieved parity errors of the E-step kernels against the fp64 oracle (numbers behind the tolerances
the tests state): emissions, posteriors, log-likelihoods, transition counts, statistics, parameters.
    python profiles/exp_parity.py            (one GPU)"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fast
from poccala_b200 import synth
from poccala_b200.engine import Corpus, Engine, EStep, Model

eng = Engine(0)
out = {}
for name, (n_utt, T, L, n_units, mix, seed) in {"ragged_T90_L5_M4": (20, 90, 5, 5, 4, 4), "T300_L10_M16": (12, 300, 10, 12, 16, 9),
                                                  "T300_L10_M64": (6, 300, 10, 12, 64, 10)}.items():
    truth, init, labels, utts = synth.make_corpus(n_utt, T, L, n_units, mix, seed, ragged=name.startswith("ragged"))
    tm = synth.default_transmat(n_units)
    om = fast.Model(*init, tm)
    corpus = Corpus(eng, labels, np.array([len(x) for x in utts], dtype=np.int32), n_units)
    res = {}
    for k2 in (1, 0):
        model = Model(eng, *init, tm)
        es = EStep(eng, corpus, model)
        es.load_frames(torch.as_tensor(np.concatenate(utts)).to(eng.device))
        eng.set_option("k2_kernel", k2)
        es.estep()
        torch.cuda.synchronize()
        eng.set_option("k2_kernel", 1)
        logp = es.utt_logp.cpu().numpy()
        pt = es.pair_trans.cpu().numpy()
        e = dict(b_rel=0.0, gamma_abs=0.0, gamma_rel_floor1e2=0.0, gamma_sum_dev=0.0, logp_rel=0.0, trans_abs=0.0)
        for u, (lab, X) in enumerate(zip(labels, utts)):
            r = fast.estep_batch(om, np.asarray(lab)[None], X[None], keep=True)
            c = fast.score_components_direct(om, np.asarray(lab)[None], X[None])
            b_ref = fast.lse(c, axis=-1)[0].T
            bb = corpus.emission_view(es.b, u).cpu().numpy()
            e["b_rel"] = max(e["b_rel"], float(np.max(np.abs(bb - b_ref) / np.abs(b_ref))))
            g = np.exp(corpus.emission_view(es.lgam, u).cpu().numpy().astype(np.float64))
            gr = np.exp(r["lgam"][0].T)
            e["gamma_abs"] = max(e["gamma_abs"], float(np.abs(g - gr).max()))
            e["gamma_rel_floor1e2"] = max(e["gamma_rel_floor1e2"], float(np.max(np.abs(g - gr) / np.maximum(gr, 1e-2))))
            e["gamma_sum_dev"] = max(e["gamma_sum_dev"], float(np.abs(g[:, 1:].sum(axis=0) - gr[:, 1:].sum(axis=0)).max()))
            e["logp_rel"] = max(e["logp_rel"], abs(logp[u] - r["logp"][0]) / abs(r["logp"][0]))
            p0 = corpus.pair_off[u]
            for p in range(len(lab)):
                for rr in range(3):
                    s = 1 + 3 * p + rr
                    ref = np.array([r["k_self"][0][s], r["k_next"][0][s], r["gamma"][0][s]])
                    got = logp[u] + pt[p0 + p, 3 * rr:3 * rr + 3].astype(np.float64)
                    fin = np.isfinite(ref)
                    assert (np.isfinite(got) == fin).all(), (name, k2, u, p, rr, got, ref)
                    if fin.any():
                        e["trans_abs"] = max(e["trans_abs"], float(np.abs(got[fin] - ref[fin]).max()))
        stats, info = fast.estep_corpus(om, labels, utts)
        acc = es.acc.cpu().numpy().reshape(n_units, 3, mix, 80)
        e["occ_rel_floor1e2"] = float(np.max(np.abs(acc[..., 39] - stats.occ) / np.maximum(stats.occ, 1e-2)))
        es.mstep(c_covariance=1e-6)
        torch.cuda.synchronize()
        new = fast.mstep(om, stats, c_covariance=1e-6)
        mean, var, alpha, tmn = model.numpy()
        ok = stats.occ >= 1e-4
        e["alpha_rel_floor1e3"] = float(np.max(np.abs(alpha - new.alpha) / np.maximum(new.alpha, 1e-3)))
        e["mean_err_over_sd"] = float(np.max((np.abs(mean - new.mean) / np.sqrt(new.var))[ok]))
        gvar = np.concatenate(utts).var(axis=0)
        e["var_rel_floor_1pct_gvar"] = float(np.max((np.abs(var - new.var) / np.maximum(new.var, 1e-2 * gvar))[ok]))
        e["var_rel"] = float(np.max((np.abs(var - new.var) / new.var)[ok]))
        e["transmat_rel_floor1e2"] = float(np.max(np.abs(tmn - new.transmat) / np.maximum(new.transmat, 1e-2)))
        res["k2_kernel=%d" % k2] = e
    out[name] = res
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/parity_errors.json", "w"), indent=1)
