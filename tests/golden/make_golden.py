"""Generate tests/golden/*.npz by EXECUTING the unmodified reference (/root/reference).

Run here (the build container) only:  python tests/golden/make_golden.py
The GPU box has no reference tree; it consumes the committed .npz files.
Recipe: SURVEY.md Appendix B via oracle/ref_harness.py.
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")

from oracle import ref_harness as rh  # noqa: E402
from poccala_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
UNITS = ["a", "b", "c"]
MIX = 4


def estep_golden():
    truth, init, _, _ = synth.make_corpus(1, 30, 3, 3, MIX, cfg_seed=1)
    mean, var, alpha = init
    label_ids = [np.array([0]), np.array([1, 2]), np.array([0, 1, 0]), np.array([2, 2, 1, 0])]
    lens = [24, 17, 40, 33]
    utts = [synth.make_utterance(l, t, truth, 4242 + i) for i, (l, t) in enumerate(zip(label_ids, lens))]
    H = rh.Harness(UNITS, MIX)
    params = {u: [(mean[i, r], var[i, r], alpha[i, r]) for r in range(3)] for i, u in enumerate(UNITS)}
    out = dict(mean=mean, var=var, alpha=alpha, n_utt=len(utts))
    for k, (lab, X) in enumerate(zip(label_ids, utts)):
        label = [UNITS[i] for i in lab]
        r = H.estep_in_memory(label, X, params)
        out[f"u{k}_label"] = lab
        out[f"u{k}_X"] = X
        for key in ("A", "B", "alpha", "beta", "ksai", "gamma", "pi_next"):
            out[f"u{k}_{key}"] = np.array(r[key])
        for p, h in enumerate(r["hmm_list"]):
            out[f"u{k}_p{p}_ksai_acc"] = np.array(h.ksai_acc)
            out[f"u{k}_p{p}_gamma_acc"] = np.array(h.gamma_acc)
            for gi, g in enumerate(h.profunction[1:-1]):
                out[f"u{k}_p{p}_g{gi}_acc"] = np.array(g.acc)
                out[f"u{k}_p{p}_g{gi}_alpha_acc"] = np.array(g.alpha_acc)
                out[f"u{k}_p{p}_g{gi}_mean_acc"] = np.array(g.mean_acc)
                out[f"u{k}_p{p}_g{gi}_cov_acc"] = np.array(g._GMM__covariance_acc)
        sc, path = H.viterbi(r["states"], r["A"], r["B"], r["pi0"])
        out[f"u{k}_vit_score"] = sc
        out[f"u{k}_vit_path"] = path
        _, lab_path = H.viterbi_labels(r["states"], r["A"], r["B"], r["pi0"])
        out[f"u{k}_vit_units"] = np.array([UNITS.index(x) for x in lab_path])
    # one full file-based EM iteration through the reference's own code path (two rounds)
    H2 = rh.Harness(UNITS, MIX)
    new = H2.file_based_iteration(params, [([UNITS[i] for i in l], X) for l, X in zip(label_ids, utts)], 1e-6)
    out["it1_transmat"] = np.stack([new[u]["transmat"] for u in UNITS])
    out["it1_mean"] = np.stack([np.stack([g["mean"] for g in new[u]["gmms"]]) for u in UNITS])
    out["it1_var"] = np.stack([np.stack([g["var"] for g in new[u]["gmms"]]) for u in UNITS])
    out["it1_alpha"] = np.stack([np.stack([g["alpha"] for g in new[u]["gmms"]]) for u in UNITS])
    np.savez_compressed(os.path.join(OUT, "estep_small.npz"), **out)
    print("estep_small.npz written")


def viterbi_ties_golden():
    """Integer-valued emissions force exact ties; the reference picks the lower state index."""
    H = rh.Harness(UNITS, MIX)
    rng = np.random.default_rng(99)
    out = {}
    for k, (L, T) in enumerate([(1, 6), (2, 12), (4, 25), (3, 1)]):
        N = 3 * L + 2
        A = np.zeros((N, N))
        A[0, 1] = 1.0
        for j in range(1, N - 1):
            A[j, j] = 0.5
            A[j, j + 1] = 0.5
        B = rng.integers(-3, 1, size=(N, T)).astype(np.float64)
        B[0] = 0.0
        B[-1] = -np.inf
        pi = np.ones(N) / N
        states = {i: "a" for i in range(N)}
        sc, path = H.viterbi(states, A, B, pi)
        out[f"v{k}_A"] = A
        out[f"v{k}_B"] = B
        out[f"v{k}_score"] = sc
        out[f"v{k}_path"] = path
    out["n"] = 4
    np.savez_compressed(os.path.join(OUT, "viterbi_ties.npz"), **out)
    print("viterbi_ties.npz written")


def kmeans_golden():
    H = rh.Harness(UNITS, MIX)
    out = {}
    cases = [(60, 2, 5, 3), (150, 4, 39, 11), (200, 5, 39, 12345), (40, 1, 39, 7)]
    for k, (n, K, D, seed) in enumerate(cases):
        rng = np.random.default_rng(seed)
        centers = rng.normal(0, 3, size=(max(K, 2), D))
        data = centers[rng.integers(0, max(K, 2), size=n)] + rng.normal(size=(n, D))
        mean, cov, alpha, clustered = H.kmeans(list(data), K, seed)
        members = []
        for cl in clustered:
            idx = []
            for p in cl:
                hit = np.where((data == p).all(axis=1))[0]
                idx.append(int(hit[0]))
            members.append(idx)
        out[f"k{k}_data"] = data
        out[f"k{k}_K"] = K
        out[f"k{k}_seed"] = seed
        out[f"k{k}_mean"] = mean
        out[f"k{k}_var"] = np.stack([np.diag(c) for c in cov])
        out[f"k{k}_alpha"] = np.array(alpha)
        out[f"k{k}_sizes"] = np.array([len(m) for m in members])
        out[f"k{k}_members"] = np.concatenate([np.array(m, dtype=np.int64) for m in members])
    out["n"] = len(cases)
    np.savez_compressed(os.path.join(OUT, "kmeans_small.npz"), **out)
    print("kmeans_small.npz written")


def flat_start_golden():
    """AcousticModel.__flat_start (AcousticModel.py:479-517) executed as is: only the audio loader is
    replaced by in-memory feature arrays and the parameter writer by a capture."""
    import random

    H = rh.Harness(UNITS, MIX)
    am = H.am
    rng = np.random.default_rng(2024)
    utts = [rng.normal(0.5, 2.0, size=(int(n), 39)) for n in rng.integers(15, 30, size=8)]
    paths = [["utt%d" % i, "lab%d" % i] for i in range(len(utts))]
    am._AcousticModel__load_audio = lambda path: utts[int(path[3:])]
    captured = {}

    def save(unit, hmm):
        gm = hmm.profunction[1:-1]
        captured[unit] = (np.stack([np.array(g.mean) for g in gm]),
                          np.stack([np.stack([np.diag(c) for c in g.covariance]) for g in gm]))

    am._AcousticModel__save_parameter = save
    am.delete_trainInfo = lambda: None
    am._AcousticModel__loaded_units = list(UNITS)
    random.seed(31)
    np.random.seed(32)
    am._AcousticModel__flat_start(paths, len(paths), proportion=0.5, step=2, differentiation=True, coefficient=0.7)
    out = dict(n_utt=len(utts), proportion=0.5, step=2, coefficient=0.7, py_seed=31, np_seed=32)
    for i, x in enumerate(utts):
        out[f"x{i}"] = x
    out["mean"] = np.stack([captured[u][0] for u in UNITS])
    out["var"] = np.stack([captured[u][1] for u in UNITS])
    np.savez_compressed(os.path.join(OUT, "flat_start.npz"), **out)
    print("flat_start.npz written", out["mean"].shape, out["var"].shape)


def gmm_em_golden():
    """Clustering.GMM.em(smem=False) (Clustering.py:583-651,695-719) executed on two small problems."""
    H = rh.Harness(UNITS, MIX)
    out = {}
    for c, (n, M, seed) in enumerate([(160, 3, 5), (240, 4, 6)]):
        rng = np.random.default_rng(seed)
        centres = rng.normal(0, 0.35, size=(M, 39))
        data = centres[rng.integers(0, M, size=n)] + rng.normal(size=(n, 39)) * rng.uniform(0.6, 1.2, size=(1, 39))
        mean0 = centres + rng.normal(0, 0.25, size=centres.shape)
        var0 = rng.uniform(0.8, 1.6, size=(M, 39))
        alpha0 = np.ones(M) / M
        g = H.Clustering.GMM(H.log, dimension=39, mix_level=M, data=list(data), alpha=alpha0.copy(), mean=mean0.copy(),
                             covariance=np.stack([np.diag(v) for v in var0]))
        iters = {"n": 0}
        orig = g.expectation

        def counted(orig=orig, iters=iters):
            iters["n"] += 1
            return orig()

        g.expectation = counted
        g.em(show_q=False, smem=False, c_covariance=1e-3)
        out[f"e{c}_data"] = data
        out[f"e{c}_mean0"] = mean0
        out[f"e{c}_var0"] = var0
        out[f"e{c}_alpha0"] = alpha0
        out[f"e{c}_mean"] = np.array(g.mean)
        out[f"e{c}_var"] = np.stack([np.diag(x) for x in g.covariance])
        out[f"e{c}_alpha"] = np.array(g.alpha)
        out[f"e{c}_iters"] = iters["n"]
    out["n"] = 2
    np.savez_compressed(os.path.join(OUT, "gmm_em.npz"), **out)
    print("gmm_em.npz written", [int(out[f"e{c}_iters"]) for c in range(2)])


def smem_golden():
    """Clustering.GMM.em(smem=True) (Clustering.py:483-577,695-719) executed on three small problems; what the
    split / merge search ranked, chose and decided is recorded next to the final parameters."""
    import random

    H = rh.Harness(UNITS, MIX)
    out = {}
    problems = [(160, 3, 0.35, 5), (240, 4, 2.0, 6), (300, 5, 1.0, 7)]
    for c, (n, M, spread, seed) in enumerate(problems):
        rng = np.random.default_rng(seed)
        centres = rng.normal(0, spread, size=(M, 39))
        data = centres[rng.integers(0, M, size=n)] + rng.normal(size=(n, 39)) * rng.uniform(0.6, 1.2, size=(1, 39))
        mean0 = centres[rng.integers(0, M, size=M)] + rng.normal(0, 0.25, size=centres.shape)
        var0 = rng.uniform(0.8, 1.6, size=(M, 39))
        alpha0 = np.ones(M) / M
        g = H.Clustering.GMM(H.log, dimension=39, mix_level=M, data=list(data), alpha=alpha0.copy(), mean=mean0.copy(),
                             covariance=np.stack([np.diag(v) for v in var0]))
        rec = {"q": [], "max": [], "exp": 0}
        merge, split, smem, qf, mx, ex = (g._GMM__J_merge, g._GMM__J_split, g._GMM__SMEM, g.q_function, g.maximization,
                                          g.expectation)

        def w_merge():
            rec["merge"] = merge()
            return rec["merge"]

        def w_split():
            rec["split"] = split()
            return rec["split"]

        def w_smem(q, **kw):
            rec["q_in"] = q
            rec["q"], rec["max"] = [], []
            # the tables the search starts from: log posteriors of the last expectation(), current parameters
            rec["gamma_in"] = np.array(g._GMM__gamma)
            rec["mean_in"], rec["alpha_in"] = np.array(g.mean), np.array(g.alpha)
            rec["var_in"] = np.stack([np.diag(x) for x in g.covariance])
            rec["ret"] = smem(q, **kw)
            return rec["ret"]

        def w_q():
            rec["q"].append(qf())
            return rec["q"][-1]

        def w_max(**kw):
            rec["max"].append(mx(**kw))
            return rec["max"][-1]

        def w_exp():
            rec["exp"] += 1
            return ex()

        g._GMM__J_merge, g._GMM__J_split, g._GMM__SMEM = w_merge, w_split, w_smem
        g.q_function, g.maximization, g.expectation = w_q, w_max, w_exp
        np.random.seed(100 + c)
        random.seed(200 + c)
        g.em(show_q=False, smem=True, c_covariance=1e-3)
        assert rec["ret"] is False, "an accepted candidate cannot be continued by the reference (mix_level stays M-3)"
        out[f"s{c}_data"], out[f"s{c}_mean0"], out[f"s{c}_var0"], out[f"s{c}_alpha0"] = data, mean0, var0, alpha0
        out[f"s{c}_merge"] = np.array([[r[0], r[1], float(np.ravel(r[2])[0])] for r in rec["merge"]])
        out[f"s{c}_split"] = np.array([[r[0], float(np.ravel(r[1])[0])] for r in rec["split"]])
        out[f"s{c}_q_in"] = float(rec["q_in"])
        out[f"s{c}_gamma_in"], out[f"s{c}_mean_in"] = rec["gamma_in"], rec["mean_in"]
        out[f"s{c}_var_in"], out[f"s{c}_alpha_in"] = rec["var_in"], rec["alpha_in"]
        out[f"s{c}_q12"] = np.array([float(v) for v in rec["q"]])  # q_1 (three new components), q_2 (the rest)
        nm, nc, na = rec["max"][-1]
        out[f"s{c}_new_mean"] = np.array(nm)
        out[f"s{c}_new_var"] = np.stack([np.diag(x) for x in nc])
        out[f"s{c}_new_alpha"] = np.array(na)
        out[f"s{c}_mean"] = np.array(g.mean)
        out[f"s{c}_var"] = np.stack([np.diag(x) for x in g.covariance])
        out[f"s{c}_alpha"] = np.array(g.alpha)
        out[f"s{c}_iters"] = rec["exp"]
        out[f"s{c}_seeds"] = np.array([100 + c, 200 + c])
        print(c, "merge", out[f"s{c}_merge"][:3].tolist(), "split", out[f"s{c}_split"].tolist(), "q", rec["q_in"], rec["q"])
    out["n"] = len(problems)
    np.savez_compressed(os.path.join(OUT, "gmm_smem.npz"), **out)
    print("gmm_smem.npz written")


def mfcc_golden():
    """AudioProcessing.MFCC.mfcc + AudioProcessing.VAD (AudioProcessing.py:146-542) executed on two synthetic wave
    files (16 kHz mono; 8 kHz stereo).  numpy 2 removed the binary mode of np.fromstring that init_audio calls
    (:164): it is mapped to np.frombuffer for the run - the same bytes -> int16 conversion."""
    import tempfile
    import wave

    rh.Harness(UNITS, MIX)
    from StatisticalModel.AudioProcessing import AudioProcessing

    np.fromstring = lambda b, dtype=float: np.frombuffer(b, dtype=dtype).copy()
    out = {}
    cases = [(16000, 1, 1.0, 3), (8000, 2, 1.5, 4)]
    for c, (sr, nch, secs, seed) in enumerate(cases):
        rng = np.random.default_rng(seed)
        n = int(sr * secs)
        t = np.arange(n) / sr
        voiced = (t > 0.3 * secs) & (t < 0.8 * secs)
        sig = 3000 * np.sin(2 * np.pi * 220 * t) * voiced + 900 * np.sin(2 * np.pi * 1330 * t + 1.0) * voiced
        chans = [(sig * g + 200 * rng.normal(size=n)).astype(np.int16) for g in ([1.0] if nch == 1 else [1.0, 0.6])]
        pcm = np.stack(chans, axis=1).reshape(-1)
        path = os.path.join(tempfile.mkdtemp(), "x.wav")
        w = wave.open(path, "wb")
        w.setnchannels(nch); w.setsampwidth(2); w.setframerate(sr); w.writeframes(pcm.tobytes()); w.close()
        m = AudioProcessing.MFCC(13)
        m.init_audio(path=path)
        feat = m.mfcc(nfft=512, d1=True, d2=True)
        v = AudioProcessing.VAD()
        v.init_mfcc(feat)
        dist = v.mel_distance()
        osf = v.osf(dist)
        kept = v.detect(osf)
        out[f"m{c}_pcm"], out[f"m{c}_rate"], out[f"m{c}_channels"] = pcm, sr, nch
        out[f"m{c}_data"] = np.array(m.data)
        out[f"m{c}_mfcc"], out[f"m{c}_dist"], out[f"m{c}_osf"], out[f"m{c}_kept"] = feat, dist, osf, kept
        out[f"m{c}_static"] = m.mfcc(nfft=512, cal_energy=False)
        print(c, "frames", feat.shape, "kept", kept.shape, "finite", bool(np.isfinite(feat).all()))
    out["n"] = len(cases)
    np.savez_compressed(os.path.join(OUT, "mfcc.npz"), **out)
    print("mfcc.npz written")


def alignment_golden():
    """Mode-1 data preparation executed as is: __eq_segment(mode='e') (AcousticModel.py:605-612), the
    post-Viterbi part of multi_process_data (:750-764, with discriminate :937-955) and __get_gmmdata
    (:630-644).  Frames are their own corpus index ([T,1] arrays), the segment writer is replaced by a
    capture, and for the aligned case the model set-up / Viterbi calls in front of line 750 are
    stubbed to hand back the unit sequence under test."""
    from unittest.mock import MagicMock

    units = ["a", "b", "c", "d"]
    H = rh.Harness(units, MIX)
    am = H.am
    am.log = rh._Quiet()  # multi_process_data closes its log after every utterance
    rng = np.random.default_rng(77)
    out = {}

    def run_case(tag, labels, lens, sequences):
        captured = {}
        am._AcousticModel__save_data = lambda unit, d: captured.setdefault(unit, []).append(np.array(d))
        off = 0
        for u, (lab, T) in enumerate(zip(labels, lens)):
            data = (np.arange(T) + off)[:, None].astype(np.float64)
            names = [units[i] for i in lab]
            if sequences is None:
                am._AcousticModel__eq_segment(data, names, mode='e')
            else:
                seq = np.array([units[i] for i in sequences[u]])
                am.init_unit = lambda unit=None, new_log=True, fix_code=0: MagicMock()
                am.init_parameter = lambda unit=None, hmm=None: None
                am.embedded = lambda *a, **k: (None, None, np.zeros((1, 1)), None)
                am.viterbi = lambda *a, seq=seq: (0., seq)
                cwd = os.getcwd()
                os.chdir(H.tmp)  # the reference drops a prob.csv into the working directory (:746)
                try:
                    am.multi_process_data(names, data, False, u + 1, len(labels), 2)
                finally:
                    os.chdir(cwd)
            off += T
        n_sets = 0
        for i, unit in enumerate(units):
            if unit not in captured:
                continue
            g = am._AcousticModel__get_gmmdata(captured[unit])
            for r in range(3):
                out[f"{tag}_set_{i * 3 + r}"] = np.asarray(g[r]).reshape(-1).astype(np.int64)
                n_sets += 1
        out[f"{tag}_n_utt"] = len(labels)
        out[f"{tag}_lens"] = np.array(lens)
        for u, lab in enumerate(labels):
            out[f"{tag}_label{u}"] = np.array(lab)
            if sequences is not None:
                out[f"{tag}_seq{u}"] = np.array(sequences[u])
        return n_sets

    # uniform: remainders dropped, chunks shorter than 3 frames, repeated units
    labels = [[0, 1, 2], [1, 1, 3, 0], [2], [0, 3, 0, 3, 1, 2, 1], [3, 2]]
    lens = [31, 50, 7, 16, 5]
    n0 = run_case("uni", labels, lens, None)

    # aligned: random monotone unit sequences; runs of 1-2 frames; adjacent equal units merge into one
    # run; one utterance stops short of its last unit (discarded), one skips nothing but repeats units
    def walk(lab, T, stop_early=False):
        L = len(lab) - (1 if stop_early else 0)
        cuts = np.sort(rng.choice(np.arange(1, T), size=L - 1, replace=False)) if L > 1 else np.array([], dtype=int)
        pos = np.searchsorted(cuts, np.arange(T), side="right")
        return [lab[p] for p in pos]

    labels = [[0, 1, 2], [1, 1, 3, 0], [2, 0, 2], [0, 3, 0, 3, 1, 2, 1], [3, 2, 1], [0, 1]]
    lens = [40, 33, 9, 64, 25, 2]
    seqs = [walk(l, T) for l, T in zip(labels, lens)]
    seqs[4] = walk(labels[4], lens[4], stop_early=True)  # unit 1 never reached -> utterance dropped
    n1 = run_case("ali", labels, lens, seqs)
    np.savez_compressed(os.path.join(OUT, "alignment.npz"), **out)
    print("alignment.npz written", n0, n1)


if __name__ == "__main__":
    assert rh.available(), "needs /root/reference"
    if "--alignment-only" in sys.argv:
        alignment_golden()
        sys.exit(0)
    if "--flat-start-only" in sys.argv:
        flat_start_golden()
        sys.exit(0)
    if "--gmm-em-only" in sys.argv:
        gmm_em_golden()
        sys.exit(0)
    if "--mfcc-only" in sys.argv:
        mfcc_golden()
        sys.exit(0)
    if "--smem-only" in sys.argv:
        smem_golden()
        sys.exit(0)
    estep_golden()
    viterbi_ties_golden()
    kmeans_golden()
    flat_start_golden()
    gmm_em_golden()
    alignment_golden()
