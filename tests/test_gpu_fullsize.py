"""BASELINE.json's full sizes through size-independent properties (the oracle needs core-hours there):
configs[1] (1 000 utterances x 300 frames x 10 units of 57, 16 mixtures) and a shard of configs[4] (the same shape
with 64 mixtures) - posterior mass is conserved
from K2 into K3's statistics, first moments match a direct fp64 reduction, utterance shards add up and
per-utterance results do not depend on the batch they ran in; configs[3] (10 000 x 1 000 frames x 20
units) - every Viterbi path is a legal monotone walk whose recomputed score equals the returned one,
and a random sample of utterances is bit-exact against the fp64 oracle."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import fast  # noqa: E402  (checker)
from poccala_b200 import synth  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from poccala_b200.engine import Engine

    e = Engine(0)
    yield e
    e.close()


def _estep(eng, init, labels, n_frames, x, n_units, shift=None, inv_scale=None):
    from poccala_b200.engine import Corpus, EStep, Model

    corpus = Corpus(eng, labels, np.asarray(n_frames, dtype=np.int32), n_units)
    model = Model(eng, *init, synth.default_transmat(n_units))
    es = EStep(eng, corpus, model)
    es.load_frames(x, shift=shift, inv_scale=inv_scale)
    es.score()
    es.forward_backward()
    es.accumulate()
    return corpus, model, es


@pytest.mark.parametrize("U,M", [(1000, 16), (1500, 64)])
def test_full_size_mass_conservation_and_shard_additivity(eng, U, M):
    """(1000, 16): configs[1] as it is; (1500, 64): the configs[4] model (64 mixtures: the kernels for wide units) on
    a 1 500-utterance shard of its corpus."""
    T, L, NU = 300, 10, 57
    truth, init, labels, utts = synth.make_corpus(U, T, L, NU, M, 2)
    x = torch.as_tensor(np.concatenate(utts, axis=0)).to(eng.device)
    corpus, model, es = _estep(eng, init, labels, [T] * U, x, NU)
    torch.cuda.synchronize()
    sp = (3 * L + 7) & ~7
    lg = es.lgam.view(U, T, sp)[:, :, :3 * L].double()
    w = lg.exp().sum(-1)  # posterior mass of the emitting states per frame
    # A.3: per-frame normalised posteriors; the entry state holds (part of) frame 0 only (Q3)
    assert (w[:, 1:] - 1).abs().max().item() < 1e-4
    assert w[:, 0].max().item() < 1 + 1e-4
    acc = es.acc.view(NU, 3, M, -1)
    occ = acc[..., 39]  # layout of a statistics row: sum g x [0:39], sum g [39], sum g x^2 [40:79]
    # K3 accounts for the mass K2 assigned, in total and per (unit, state), within the 1e-4 bound of
    # the path: its component posteriors are exp(c - b) with c recomputed by its own contraction (four
    # fp16 partial products) and b from K1 (three), a ~1.5e-5 relative difference in sum_m exp(c - b)
    MASS = 1e-4
    assert abs(occ.sum().item() - w.sum().item()) < MASS * w.sum().item()
    state_mass = lg.exp().sum(1).view(U, L, 3)  # [U, position, state]
    lab = torch.as_tensor(np.stack(labels).astype(np.int64)).to(eng.device)
    want = torch.zeros(NU, 3, dtype=torch.float64, device=eng.device).index_add_(0, lab.view(-1), state_mass.view(-1, 3))
    got = occ.sum(-1)
    assert ((got - want).abs() <= MASS * want.clamp_min(1.0)).all()
    # first moments in the standardised coordinates the accumulators live in
    xs = (x.double() - es.shift[:39]) * es.inv_scale[:39]
    want_sx = (w.view(-1, 1) * xs).sum(0)
    got_sx = acc[..., :39].sum(dim=(0, 1, 2))
    assert ((got_sx - want_sx).abs() <= MASS * (w.view(-1, 1) * xs.abs()).sum(0)).all()
    want_sxx = (w.view(-1, 1) * xs * xs).sum(0)
    got_sxx = acc[..., 40:79].sum(dim=(0, 1, 2))
    assert ((got_sxx - want_sxx).abs() <= MASS * want_sxx).all()
    # shards: even / odd utterances with the corpus-wide standardisation
    total = torch.zeros_like(es.acc)
    for rank in range(2):
        idx = list(range(rank, U, 2))
        xr = x.view(U, T, 39)[idx].reshape(-1, 39).contiguous()
        c2, m2, e2 = _estep(eng, init, [labels[i] for i in idx], [T] * len(idx), xr, NU, es.shift, es.inv_scale)
        torch.cuda.synchronize()
        # an utterance's likelihood, iteration count and posteriors do not depend on its batch
        assert torch.equal(e2.utt_logp, es.utt_logp[idx])
        assert torch.equal(e2.utt_iters, es.utt_iters[idx])
        assert torch.equal(e2.lgam.view(len(idx), T, sp)[:, :, :3 * L], es.lgam.view(U, T, sp)[idx][:, :, :3 * L])
        total += e2.acc
    scale = es.acc.abs().amax(dim=1, keepdim=True).clamp_min(1e-3)
    assert ((total - es.acc).abs() <= 1e-5 * scale).all()


def test_cfg4_full_size_viterbi_paths_are_legal_and_scores_recompute(eng):
    from poccala_b200.engine import host_log_bands, viterbi

    U0, REP, T, L, NU, M = 500, 20, 1000, 20, 57, 16
    truth, init, labels0, utts = synth.make_corpus(U0, T, L, NU, M, 4, n_initials=22)
    U = U0 * REP
    x0 = torch.as_tensor(np.concatenate(utts, axis=0).astype(np.float32)).to(eng.device)
    gen = torch.Generator(device=eng.device).manual_seed(4)
    x = x0.repeat(REP, 1)
    x += 0.05 * torch.randn(x.shape, generator=gen, device=eng.device)  # 10 000 distinct utterances
    labels = labels0 * REP
    from poccala_b200.engine import Corpus, EStep, Model

    corpus = Corpus(eng, labels, np.full(U, T, dtype=np.int32), NU)
    tm = synth.default_transmat(NU)
    tm[:, 1:4, :] = 0
    rng = np.random.default_rng(4)
    stay = rng.uniform(0.3, 0.9, size=(NU, 3))
    for r in range(3):  # trained-looking transitions: stay / advance differ per state
        tm[:, r + 1, r + 1] = stay[:, r]
        tm[:, r + 1, r + 2] = 1 - stay[:, r]
    model = Model(eng, *init, tm)
    es = EStep(eng, corpus, model)
    es.load_frames(x)
    es.score()
    ls, ln = host_log_bands(tm, eng.device)
    N = 3 * L + 2
    logpi = torch.full((U,), float(np.log(np.ones(N) / N)[0]), dtype=torch.float64, device=eng.device)
    score, path, units = viterbi(eng, corpus, es.b, ls, ln, utt_logpi=logpi)
    torch.cuda.synchronize()
    P = path.view(U, T).long()
    d = P[:, 1:] - P[:, :-1]
    assert ((d == 0) | (d == 1)).all() and (P[:, 0] <= 1).all() and (P >= 0).all() and (P < N).all()
    lab = torch.as_tensor(np.stack(labels).astype(np.int64)).to(eng.device)  # [U, L]
    pos = ((P - 1).clamp_min(0) // 3).clamp_max(L - 1)
    assert torch.equal(units.view(U, T).long(), lab.gather(1, pos))
    # recompute the score along the returned path: log pi + emissions + transitions (fp64)
    sp = (3 * L + 7) & ~7
    B = es.b.view(U, T, sp)
    emit = B.gather(2, (P - 1).clamp(0, 3 * L - 1).unsqueeze(-1)).squeeze(-1).double()
    emit = torch.where(P == 0, torch.zeros_like(emit), emit)
    assert (P < N - 1).all()  # the exit state (emission -inf) is never on a best path
    local = torch.where(P == 0, torch.zeros_like(P), (P - 1) % 3 + 1)  # state inside the unit's 5-state HMM
    unit = lab.gather(1, pos)
    stay_lp = ls[unit[:, :-1], local[:, :-1]]
    move_lp = ln[unit[:, :-1], local[:, :-1]]
    trans = torch.where(d == 0, stay_lp, move_lp)
    want = logpi + emit.sum(1) + trans.sum(1)
    assert ((score - want).abs() <= 1e-10 * want.abs()).all()
    # a random sample against the fp64 oracle: score and path bit-exact
    om = fast.Model(*init, tm)
    for u in rng.choice(U, size=48, replace=False):
        e = corpus.emission_view(es.b, int(u)).cpu().numpy().astype(np.float64)
        ols, oln = fast.banded_transitions(om, np.asarray(labels[u])[None])
        sc, pa = fast.viterbi_banded(ols, oln, fast.full_emissions(e.T[None]))
        assert score[u].item() == sc[0]
        assert (P[u].cpu().numpy() == pa[0]).all()
