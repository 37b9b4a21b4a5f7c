"""BASELINE.json configs at sizes the oracle finishes in seconds, plus size-independent properties:
cfg 1 (single 3-state HMM, 4-mix, 100 x 300 frames) in full; cfg 4 shape (T = 1000, N = 62) Viterbi
bit-exact; cfg 5 pipeline (uniform segmentation -> k-means init -> EM iterations, utterances sharded
over two ranks) with the two shards' statistics reduced exactly as the NCCL path reduces them."""
import random

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import fast  # noqa: E402  (checker)
from poccala_b200 import synth  # noqa: E402

pytestmark = pytest.mark.gpu
REL = 1e-4
OCC_MIN = 1e-4


@pytest.fixture(scope="module")
def eng():
    from poccala_b200.engine import Engine

    e = Engine(0)
    yield e
    e.close()


def _estep(eng, init, labels, utts, n_units, tm=None):
    from poccala_b200.engine import Corpus, EStep, Model

    tm = synth.default_transmat(n_units) if tm is None else tm
    corpus = Corpus(eng, labels, np.array([len(x) for x in utts], dtype=np.int32), n_units)
    model = Model(eng, *init, tm)
    es = EStep(eng, corpus, model)
    es.load_frames(torch.as_tensor(np.concatenate(utts, axis=0)).to(eng.device))
    return corpus, model, es


def _check_model(model, new, stats, utts):
    mean, var, alpha, tm = model.numpy()
    assert np.all(np.abs(alpha - new.alpha) <= REL * np.maximum(new.alpha, 1e-3))
    ok = stats.occ >= OCC_MIN
    assert np.all(np.abs(mean - new.mean)[ok] <= (REL * np.maximum(np.abs(new.mean), np.sqrt(new.var)))[ok])
    gvar = np.concatenate(utts, axis=0).var(axis=0)
    assert np.all((np.abs(var - new.var) <= REL * np.maximum(new.var, 1e-2 * gvar))[ok])
    assert np.all(np.abs(tm - new.transmat) <= REL * np.maximum(new.transmat, 1e-2))


def test_cfg1_single_hmm_full_size(eng):
    """configs[0]: one 3-state left-to-right HMM, 4-mix diagonal GMMs, 39-dim, 100 utterances x 300
    frames, one Baum-Welch iteration - the reference's own CPU-runnable case, in full."""
    truth, init, labels, utts = synth.make_corpus(100, 300, 1, 1, 4, 1)
    corpus, model, es = _estep(eng, init, labels, utts, 1)
    es.em_iteration(c_covariance=1e-6)
    torch.cuda.synchronize()
    om = fast.Model(*init, synth.default_transmat(1))
    stats, info = fast.estep_corpus(om, labels, utts)
    logp = es.utt_logp.cpu().numpy()
    assert np.all(np.abs(logp - info["logp"]) <= 1e-5 * np.abs(info["logp"]))
    assert (es.utt_iters.cpu().numpy() == info["iters"]).all()
    _check_model(model, fast.mstep(om, stats, c_covariance=1e-6), stats, utts)


def test_cfg4_shape_viterbi_bit_exact(eng):
    """configs[3] shape: 1000-frame utterances against 20 concatenated phone HMMs (N = 62 states, two
    per lane): scores and paths bit-exact with the fp64 oracle given identical emissions, and every
    path is a monotone walk that starts in the first two states."""
    from poccala_b200.engine import host_log_bands, viterbi

    truth, init, labels, utts = synth.make_corpus(24, 1000, 20, 57, 16, 4, n_initials=22)
    corpus, model, es = _estep(eng, init, labels, utts, 57)
    es.score()
    om = fast.Model(*init, synth.default_transmat(57))
    ls, ln = host_log_bands(om.transmat, eng.device)
    logpi = torch.as_tensor(np.full(len(utts), np.log(np.ones(62) / 62)[0])).to(eng.device)
    score, path, units = viterbi(eng, corpus, es.b, ls, ln, utt_logpi=logpi)
    torch.cuda.synchronize()
    score, path = score.cpu().numpy(), path.cpu().numpy()
    for u, lab in enumerate(labels):
        e = corpus.emission_view(es.b, u).cpu().numpy().astype(np.float64)
        ols, oln = fast.banded_transitions(om, np.asarray(lab)[None])
        sc, pa = fast.viterbi_banded(ols, oln, fast.full_emissions(e.T[None]))
        f0, f1 = corpus.frame_off[u], corpus.frame_off[u + 1]
        assert score[u] == sc[0]
        assert (path[f0:f1] == pa[0]).all()
        d = np.diff(path[f0:f1])
        assert ((d == 0) | (d == 1)).all() and path[f0] <= 1


def _uniform_state_data(labels, utts, n_units):
    """Mode-1 initialisation data (AcousticModel.py:605-612 'e' split per unit, :613-626 'g' split per
    state): utterance frames cut uniformly over its units, each chunk uniformly over 3 states."""
    per_state = [[[] for _ in range(3)] for _ in range(n_units)]
    for lab, X in zip(labels, utts):
        chunk = len(X) // len(lab)
        for p, unit in enumerate(lab):
            seg = X[p * chunk:(p + 1) * chunk]
            c = len(seg) // 3
            per_state[unit][0].append(seg[:c])
            per_state[unit][1].append(seg[c:2 * c])
            per_state[unit][2].append(seg[2 * c:])
    return [[np.concatenate(s, axis=0) for s in unit] for unit in per_state]


def test_cfg5_pipeline_kmeans_init_then_sharded_em(eng):
    """configs[4] pipeline at test size: uniform segmentation -> per-state k-means
    (ClusterInitialization.kmeans(algorithm=1), bit-exact against the oracle) -> 2 EM iterations with
    the utterances dealt to two ranks whose statistics are reduced the way
    poccala_b200.distributed reduces them (MAX of the transition maxima, then one flat SUM): the
    sharded run must reproduce the single-shard run and the oracle."""
    from poccala_b200.engine import Corpus, EStep, Model, kmeans_run, kmeans_seed_points

    n_units, mix = 4, 8
    truth, _, labels, utts = synth.make_corpus(48, 150, 4, n_units, mix, 5)
    data = _uniform_state_data(labels, utts, n_units)
    # ---- k-means init, all 12 states in one launch
    flat, seeds, off = [], [], [0]
    for u in range(n_units):
        for r in range(3):
            x = data[u][r]
            rnd = random.Random(100 * u + r)
            seeds.append(kmeans_seed_points(np.ascontiguousarray(x[:, 0]), mix, rnd))
            flat.append(x)
            off.append(off[-1] + len(x))
    out = kmeans_run(eng, torch.as_tensor(np.concatenate(flat)).to(eng.device), np.array(off), mix,
                     np.array(seeds, dtype=np.int32))
    torch.cuda.synchronize()
    mean = out["mean"].cpu().numpy().reshape(n_units, 3, mix, 39)
    var = out["var"].cpu().numpy().reshape(n_units, 3, mix, 39)
    alpha = out["alpha"].cpu().numpy().reshape(n_units, 3, mix)
    for i, (u, r) in enumerate([(u, r) for u in range(n_units) for r in range(3)][:3]):  # oracle is O(N^2): 3 states
        ref = fast.kmeans_compat(flat[i], mix, random.Random(100 * u + r))
        assert np.abs(mean[u, r] - ref["mean"]).max() == 0
        assert np.allclose(var[u, r], ref["var"], rtol=1e-14, atol=0)
    alpha = alpha / alpha.sum(-1, keepdims=True)  # Q11: seed points are counted twice, weights sum > 1
    init = (mean, np.maximum(var, 1e-4), alpha)
    # ---- EM: one shard vs two shards on the same device
    tm = synth.default_transmat(n_units)
    om = fast.Model(*init, tm)
    from poccala_b200 import _native as nat
    from poccala_b200.engine import _p, _stream

    c1, m1, e1 = _estep(eng, init, labels, utts, n_units)
    shards = []
    for rank in range(2):
        idx = list(range(rank, len(utts), 2))
        corpus = Corpus(eng, [labels[i] for i in idx], np.array([len(utts[i]) for i in idx], dtype=np.int32), n_units)
        model = Model(eng, *init, tm)
        es = EStep(eng, corpus, model)
        # what load_frames(group=...) arrives at after allreducing the moments: the corpus-wide map
        es.load_frames(torch.as_tensor(np.concatenate([utts[i] for i in idx])).to(eng.device),
                       shift=e1.shift, inv_scale=e1.inv_scale)
        shards.append((corpus, model, es))
    for it in range(2):
        e1.em_iteration(c_covariance=1e-6)
        # both "ranks": local statistics, then the two collectives spelled out with tensor ops
        for c, m, e in shards:
            e.score(); e.forward_backward(); e.accumulate()
            e.tmax.fill_(float("-inf"))
            nat.call("pc_transitions_max", eng.h, c.c, _p(e.utt_logp), _p(e.pair_trans), _p(e.tmax), _stream())
        gmax = torch.maximum(shards[0][2].tmax, shards[1][2].tmax)  # all_reduce(MAX)
        for c, m, e in shards:
            e.tmax.copy_(gmax)
            e.tsum.zero_()
            nat.call("pc_transitions_sum", eng.h, c.c, _p(e.utt_logp), _p(e.pair_trans), _p(e.tmax), _p(e.tsum), _stream())
        total = shards[0][2].flat + shards[1][2].flat  # all_reduce(SUM) of the flat buffer
        for c, m, e in shards:
            e.flat.copy_(total)
            e.mstep(c_covariance=1e-6)
        torch.cuda.synchronize()
        stats, info = fast.estep_corpus(om, labels, utts)
        new = fast.mstep(om, stats, c_covariance=1e-6)
        _check_model(m1, new, stats, utts)
        for c, m, e in shards:
            _check_model(m, new, stats, utts)  # sharded == oracle == single shard, within the tolerance
        for a, b in zip(shards[0][1].numpy(), shards[1][1].numpy()):
            assert np.array_equal(a, b)  # replicas identical: no broadcast needed
        om = new
        # every iteration is checked from identical inputs: parameters that agree to 1e-4 do not give
        # statistics that agree to 1e-4 for components holding a thousandth of a frame
        for mdl in [m1] + [m for _, m, _ in shards]:
            mdl.mean.copy_(torch.as_tensor(new.mean).to(eng.device))
            mdl.var.copy_(torch.as_tensor(new.var).to(eng.device))
            mdl.alpha.copy_(torch.as_tensor(new.alpha).to(eng.device))
            mdl.transmat.copy_(torch.as_tensor(new.transmat).to(eng.device))
