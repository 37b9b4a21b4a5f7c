"""Alignment post-processing on the device (__eq_segment, discriminate, the post-Viterbi
part of multi_process_data, __get_gmmdata) against the executed reference (tests/golden/alignment.npz)
and against the oracle restatement on seeded corpora; the grouped data feeds k-means unchanged."""
import random

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import fast  # noqa: E402  (checker)
from poccala_b200 import synth  # noqa: E402
from tests.helpers import alignment_case, load_golden  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from poccala_b200.engine import Engine

    e = Engine(0)
    yield e
    e.close()


def _keys_from_sets(sets, n_frames):
    key = np.full(n_frames, -1, dtype=np.int32)
    for k, frames in enumerate(sets):
        key[frames] = k
    return key


def _path_from_units(label, seq, rng=None):
    """A composite-state path (entry state 0, emitting states 1..3L, exit 3L+1) whose unit sequence
    is `seq`: positions advance whenever the unit changes; the emitting state inside a position
    is arbitrary (the post-processing only looks at units)."""
    pos, path = 0, []
    for t, u in enumerate(seq):
        while label[pos] != u:
            pos += 1
        r = 0 if rng is None else int(rng.integers(0, 3))
        path.append(1 + 3 * pos + r)
    return np.array(path, dtype=np.int32)


def _corpus(eng, labels, lens, n_units):
    from poccala_b200.engine import Corpus

    return Corpus(eng, [np.asarray(l, dtype=np.int32) for l in labels], np.asarray(lens, dtype=np.int32), n_units)


@pytest.mark.parametrize("tag", ["uni", "ali"])
def test_segment_keys_match_reference_golden(eng, tag):
    from poccala_b200.engine import group_frames, segment_keys

    labels, lens, seqs, sets = alignment_case(load_golden("alignment.npz"), tag)
    F = int(lens.sum())
    corpus = _corpus(eng, labels, lens, 4)
    path = None
    if seqs is not None:
        paths = [_path_from_units(list(l), s) for l, s in zip(labels, seqs)]
        paths[0][0] = 0  # the entry state belongs to the first unit,
        paths[3][-1] = 3 * len(labels[3]) + 1  # the exit state to the last (AcousticModel.py:968-976)
        path = torch.as_tensor(np.concatenate(paths)).to(eng.device)
    key, kept = segment_keys(eng, corpus, path)
    out = group_frames(eng, key, 12, torch.arange(F, dtype=torch.float64, device=eng.device)[:, None].contiguous())
    torch.cuda.synchronize()
    kept = kept.cpu().numpy()
    assert kept.tolist() == ([1] * len(lens) if seqs is None else [1, 1, 1, 1, 0, 1])
    off = out["key_off"]
    data = out["data"].cpu().numpy()[:, 0].astype(np.int64)
    for k in range(12):
        want = sets.get(k, np.zeros(0, dtype=np.int64))
        got = data[off[k]:off[k + 1]]
        assert np.array_equal(got, want if seqs is None else np.sort(want)), k


def test_segment_and_group_random_corpus_against_oracle(eng):
    """400 ragged utterances (T 1..400, 1..10 units of 57, repeated and adjacent-equal units, 10 %
    of the alignments failing): keys, kept flags, offsets, order and gathered rows."""
    from poccala_b200.engine import group_frames, segment_keys

    rng = np.random.default_rng(5)
    n_units = 57
    labels, lens, seqs = [], [], []
    for u in range(400):
        L = int(rng.integers(1, 11))
        lab = rng.integers(0, n_units if u % 3 else 4, size=L)
        T = int(rng.integers(max(L, 1), 401)) if u % 7 else L  # some utterances with one frame per unit
        stop = L - 1 if (u % 10 == 9 and L > 1) else L
        cuts = np.sort(rng.choice(np.arange(1, T), size=stop - 1, replace=False)) if stop > 1 else np.zeros(0, dtype=int)
        pos = np.searchsorted(cuts, np.arange(T), side="right")
        labels.append(lab.astype(np.int32))
        lens.append(T)
        seqs.append(lab[pos])
    off = np.concatenate([[0], np.cumsum(lens)])
    F = int(off[-1])
    corpus = _corpus(eng, labels, lens, n_units)
    x = torch.as_tensor(rng.normal(size=(F, 39))).to(eng.device)
    for mode in (0, 1):
        if mode == 0:
            segs = [fast.segment_uniform(list(l), T) for l, T in zip(labels, lens)]
            path = None
        else:
            segs = [fast.segment_alignment(list(l), s) for l, s in zip(labels, seqs)]
            path = torch.as_tensor(np.concatenate([_path_from_units(list(l), s, rng) for l, s in zip(labels, seqs)])).to(eng.device)
        want_sets = fast.state_frames(segs, off, n_units)
        want_key = _keys_from_sets(want_sets, F)
        key, kept = segment_keys(eng, corpus, path)
        out = group_frames(eng, key, n_units * 3, x)
        torch.cuda.synchronize()
        assert np.array_equal(key.cpu().numpy(), want_key)
        assert np.array_equal(kept.cpu().numpy(), np.array([0 if s is None else 1 for s in segs]))
        koff = out["key_off"]
        order = out["order"].cpu().numpy()
        assert np.array_equal(np.sort(order), np.arange(F))  # a permutation: dropped frames sort last
        n_kept = int(koff[-1])
        assert n_kept == int((want_key >= 0).sum())
        for k in range(n_units * 3):
            assert np.array_equal(order[koff[k]:koff[k + 1]], want_sets[k]), (mode, k)
        assert torch.equal(out["data"], x[torch.as_tensor(order[:n_kept].astype(np.int64)).to(eng.device)])
        if mode == 1:
            assert (kept.cpu().numpy() == 0).sum() > 10


def test_group_frames_large_and_edge_cases(eng):
    """Stable order across many sort blocks (1.2 M frames, 549 keys), every frame dropped, one frame."""
    from poccala_b200.engine import group_frames

    rng = np.random.default_rng(6)
    n_keys = 549
    key = rng.integers(-1, n_keys, size=1_200_003).astype(np.int32)
    out = group_frames(eng, torch.as_tensor(key).to(eng.device), n_keys)
    torch.cuda.synchronize()
    want = np.argsort(np.where(key < 0, n_keys, key), kind="stable")
    assert np.array_equal(out["order"].cpu().numpy(), want)
    assert np.array_equal(out["key_off"], np.concatenate([[0], np.cumsum(np.bincount(key[key >= 0], minlength=n_keys))]))
    none = group_frames(eng, torch.full((1000,), -1, dtype=torch.int32, device=eng.device), 3,
                        torch.zeros(1000, 39, device=eng.device))
    assert none["key_off"].tolist() == [0, 0, 0, 0] and none["data"].shape == (0, 39)
    one = group_frames(eng, torch.tensor([2], dtype=torch.int32, device=eng.device), 3)
    assert one["key_off"].tolist() == [0, 0, 0, 1] and one["order"].tolist() == [0]
    with pytest.raises(ValueError):
        group_frames(eng, torch.zeros(4, dtype=torch.int32, device=eng.device), 20000)


def test_resegmentation_after_viterbi_feeds_kmeans(eng):
    """Mode-1 re-estimation data flow on the device: score -> Viterbi -> segment -> group -> per-state
    k-means, each stage checked against the oracle given the previous stage's output."""
    from poccala_b200.engine import Corpus, EStep, Model, group_frames, host_log_bands, kmeans_run, \
        kmeans_seed_points, segment_keys, viterbi

    n_units, mix = 4, 4
    truth, init, labels, utts = synth.make_corpus(40, 120, 4, n_units, mix, 9)
    corpus = Corpus(eng, labels, np.array([len(x) for x in utts], dtype=np.int32), n_units)
    model = Model(eng, *init, synth.default_transmat(n_units))
    es = EStep(eng, corpus, model)
    x = torch.as_tensor(np.concatenate(utts, axis=0)).to(eng.device)
    es.load_frames(x)
    es.score()
    ls, ln = host_log_bands(synth.default_transmat(n_units), eng.device)
    logpi = torch.as_tensor(np.array([np.log(np.ones(3 * len(l) + 2) / (3 * len(l) + 2))[0] for l in labels])).to(eng.device)
    score, path, units = viterbi(eng, corpus, es.b, ls, ln, utt_logpi=logpi)
    key, kept = segment_keys(eng, corpus, path)
    out = group_frames(eng, key, n_units * 3, x)
    torch.cuda.synchronize()
    units = units.cpu().numpy()
    off = corpus.frame_off
    segs = [fast.segment_alignment([int(v) for v in lab], units[off[u]:off[u + 1]]) for u, lab in enumerate(labels)]
    want_sets = fast.state_frames(segs, off, n_units)
    order = out["order"].cpu().numpy()
    koff = out["key_off"]
    for k in range(n_units * 3):
        assert np.array_equal(order[koff[k]:koff[k + 1]], want_sets[k]), k
    assert kept.cpu().numpy().sum() >= 30
    # the grouped rows are exactly the per-state problems k-means takes
    seeds = [kmeans_seed_points(np.ascontiguousarray(out["data"][koff[k]:koff[k + 1], 0].cpu().numpy()), mix,
                                random.Random(k)) for k in range(n_units * 3)]
    km = kmeans_run(eng, out["data"], koff, mix, np.array(seeds, dtype=np.int32))
    torch.cuda.synchronize()
    xs = np.concatenate(utts, axis=0)
    for k in (0, 5, 11):
        ref = fast.kmeans_compat(xs[want_sets[k]], mix, random.Random(k))
        assert np.abs(km["mean"][k].cpu().numpy() - ref["mean"]).max() == 0
