"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the golden
vectors produced by the executed reference.  Tolerances are the ones BASELINE.json's north_star
states: Viterbi paths bit-exact given identical emissions; log-likelihoods, posteriors and
re-estimated parameters within 1e-4 relative (fp32 kernels vs the fp64 reference)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import fast  # noqa: E402  (test infrastructure: the checker)
from poccala_b200 import synth  # noqa: E402
from tests.helpers import load_golden  # noqa: E402

pytestmark = pytest.mark.gpu

REL = 1e-4


@pytest.fixture(scope="module")
def eng():
    from poccala_b200.engine import Engine

    e = Engine(0)
    yield e
    e.close()


def _corpus(eng, labels, utts, n_units):
    from poccala_b200.engine import Corpus

    return Corpus(eng, labels, np.array([len(x) for x in utts], dtype=np.int32), n_units)


def _setup(eng, init, labels, utts, n_units, transmat=None, standardise=True):
    from poccala_b200.engine import EStep, Model

    mean, var, alpha = init
    tm = synth.default_transmat(n_units) if transmat is None else transmat
    corpus = _corpus(eng, labels, utts, n_units)
    model = Model(eng, mean, var, alpha, tm)
    es = EStep(eng, corpus, model, standardise=standardise)
    es.load_frames(torch.as_tensor(np.concatenate(utts, axis=0)).to(eng.device))
    return corpus, model, es, fast.Model(mean, var, alpha, tm)


def _relerr(a, b, floor=1e-6):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


def _var_close(var, ref, utts):
    """Variances come out of an fp32 contraction as a difference of moments (sxx - 2 mu sx + mu^2 occ),
    so their absolute error scales with the spread of the data, not with the variance itself:
    1e-4 relative for every variance above 1% of the feature's global variance, and the same
    absolute error (1e-6 of the global variance) for collapsed components below that."""
    gvar = np.concatenate(utts, axis=0).var(axis=0)
    return np.all(np.abs(var - ref) <= REL * np.maximum(ref, 1e-2 * gvar))


OCC_MIN = 1e-4  # frames; see DESIGN.md "tolerances": below this a component's mean/variance are
# determined by posteriors < 1e-9 that the fp16-split tensor-core path does not resolve; its
# weight (alpha -> 0) is still compared.


def _ragged(cfg_seed=3, n_utt=14, T=60, L=4, n_units=5, mix=4):
    truth, init, labels, utts = synth.make_corpus(n_utt, T, L, n_units, mix, cfg_seed, ragged=True)
    return init, labels, utts, n_units


@pytest.mark.parametrize("standardise", [True, False])
def test_scores_match_oracle(eng, standardise):
    init, labels, utts, n_units = _ragged()
    corpus, model, es, om = _setup(eng, init, labels, utts, n_units, standardise=standardise)
    es.score()
    torch.cuda.synchronize()
    for u, (lab, X) in enumerate(zip(labels, utts)):
        c = fast.score_components_direct(om, lab[None], X[None])
        b_ref = fast.lse(c, axis=-1)[0].T  # [3L, T]
        b_gpu = corpus.emission_view(es.b, u).cpu().numpy()
        assert _relerr(b_gpu, b_ref) < REL
        # what the kernels actually reach (fp32 contraction): two orders better than required
        assert np.abs(b_gpu - b_ref).max() < 2e-3


@pytest.mark.parametrize("mix", [4, 8, 16, 32, 64])
def test_scores_tensor_core_vs_cuda_core_and_oracle(eng, mix):
    """K1 on tcgen05 (fp16 hi/lo split, 3 products) against the fp32 CUDA-core kernel and the fp64
    oracle; utterances longer than one 128-frame tile, ragged tails, many tiles per work item."""
    truth, init, labels, utts = synth.make_corpus(12, 300, 4, 5, mix, 30 + mix, ragged=True)
    utts[0] = utts[0][:1]
    corpus, model, es, om = _setup(eng, init, labels, utts, 5)
    out = {}
    for tc in (0, 1, 2):  # CUDA cores / tcgen05 (wide accumulators up to 16 mixtures) / tcgen05 per position pair
        eng.set_option("tensor_core", min(tc, 1))
        eng.set_option("k1_kernel", 0 if tc == 2 else 1)
        es.b.fill_(float("nan"))
        es.score()
        torch.cuda.synchronize()
        out[tc] = es.b.clone()
    eng.set_option("tensor_core", 1)
    eng.set_option("k1_kernel", 1)
    for u in range(len(utts)):  # both tensor-core kernels run the same contraction
        a1, a2 = corpus.emission_view(out[1], u).cpu().numpy(), corpus.emission_view(out[2], u).cpu().numpy()
        assert np.all(np.abs(a1 - a2) <= 1e-4 + 2e-7 * np.abs(a2))  # a few fp32 ulps: the FMA-pipe exponential
    worst = 0.0
    for u, (lab, X) in enumerate(zip(labels, utts)):
        c = fast.score_components_direct(om, lab[None], X[None])
        b_ref = fast.lse(c, axis=-1)[0].T
        b_tc = corpus.emission_view(out[1], u).cpu().numpy()
        b_cc = corpus.emission_view(out[0], u).cpu().numpy()
        assert np.isfinite(b_tc).all()
        assert _relerr(b_tc, b_ref) < REL
        worst = max(worst, np.abs(b_tc - b_ref).max(), np.abs(b_cc - b_ref).max())
    assert worst < 2e-3, worst


@pytest.mark.parametrize("mix", [16, 32, 64])
def test_scores_tile_and_label_shapes(eng, mix):
    """The utterance-major scoring kernels hold their frame tiles in groups of three and hand unit images through
    rings whose turn-around depends on the number of tiles and label positions: utterances of 1 .. 6 tiles (one
    frame, one tile, the tile boundaries +-1, two groups) with 1, 2 and 7 label positions, both kernel
    generations against each other and against the fp64 oracle."""
    n_units = 5
    rng = np.random.default_rng(100 + mix)
    truth = synth.make_truth(n_units, mix, 300 + mix)
    init = synth.perturb(*truth, seed=400 + mix)
    lengths = [1, 2, 127, 128, 129, 256, 257, 300, 384, 385, 513, 700]
    labels, utts = [], []
    for i, T in enumerate(lengths):
        L = (1, 2, 7)[i % 3]
        lab = rng.integers(0, n_units, size=L).astype(np.int32)
        labels.append(lab)
        utts.append(synth.make_utterance(lab, T, truth, 500 + i))
    corpus, model, es, om = _setup(eng, init, labels, utts, n_units)
    out = {}
    for k1 in (1, 0):
        eng.set_option("k1_kernel", k1)
        es.b.fill_(float("nan"))
        es.score()
        torch.cuda.synchronize()
        out[k1] = es.b.clone()
    eng.set_option("k1_kernel", 1)
    for u, (lab, X) in enumerate(zip(labels, utts)):
        c = fast.score_components_direct(om, lab[None], X[None])
        b_ref = fast.lse(c, axis=-1)[0].T
        b1 = corpus.emission_view(out[1], u).cpu().numpy()
        b0 = corpus.emission_view(out[0], u).cpu().numpy()
        assert np.isfinite(b1).all() and b1.shape == b_ref.shape
        assert np.all(np.abs(b1 - b0) <= 1e-4 + 2e-7 * np.abs(b0)), (u, len(X), len(lab))
        assert _relerr(b1, b_ref) < REL, (u, len(X), len(lab))


def test_scores_tensor_core_dead_and_scaled_rows(eng):
    """alpha = 0 (log 0 constant) and a collapsed component whose weights exceed the fp16 range."""
    truth, init, labels, utts = synth.make_corpus(6, 200, 3, 3, 16, 77)
    mean, var, alpha = [a.copy() for a in init]
    alpha[0, 0, 3] = 0.0
    alpha[0, 0] /= alpha[0, 0].sum()
    var[1, 2, 5, :] = 1e-5  # w = mu/var ~ 1e5 > fp16 max
    corpus, model, es, om = _setup(eng, (mean, var, alpha), labels, utts, 3)
    es.score()
    torch.cuda.synchronize()
    for u, (lab, X) in enumerate(zip(labels, utts)):
        c = fast.score_components_direct(om, lab[None], X[None])
        b_ref = fast.lse(c, axis=-1)[0].T
        b_tc = corpus.emission_view(es.b, u).cpu().numpy()
        assert _relerr(b_tc, b_ref) < REL


@pytest.mark.parametrize("k2_kernel,L", [(1, 5), (0, 5), (1, 14), (0, 14)])
def test_forward_backward_matches_oracle(eng, k2_kernel, L):
    """Both forward-backward kernels (one warp per utterance with the likelihood as normaliser; three
    warps with a per-frame normaliser), one and two states per lane."""
    init, labels, utts, n_units = _ragged(cfg_seed=4, n_utt=20, T=90, L=L)
    corpus, model, es, om = _setup(eng, init, labels, utts, n_units)
    es.score()
    eng.set_option("k2_kernel", k2_kernel)
    try:
        es.forward_backward()
        torch.cuda.synchronize()
    finally:
        eng.set_option("k2_kernel", 1)
    logp = es.utt_logp.cpu().numpy()
    iters = es.utt_iters.cpu().numpy()
    pt = es.pair_trans.cpu().numpy()
    for u, (lab, X) in enumerate(zip(labels, utts)):
        r = fast.estep_batch(om, lab[None], X[None], keep=True)
        assert abs(logp[u] - r["logp"][0]) <= REL * abs(r["logp"][0])
        assert abs(logp[u] - r["logp"][0]) < 5e-3  # achieved: ~1e-7 relative
        assert iters[u] == r["iters"][0]
        g_gpu = np.exp(corpus.emission_view(es.lgam, u).cpu().numpy().astype(np.float64))
        g_ref = np.exp(r["lgam"][0].T)
        # posteriors: 1e-4 relative above 5% occupancy, the same absolute error (5e-6) below
        assert np.all(np.abs(g_gpu - g_ref) <= REL * np.maximum(g_ref, 5e-2))
        L = len(lab)
        p0 = corpus.pair_off[u]
        for p in range(L):
            for rr in range(3):
                s = 1 + 3 * p + rr
                ref = np.array([r["k_self"][0][s], r["k_next"][0][s], r["gamma"][0][s]])
                got = logp[u] + pt[p0 + p, 3 * rr:3 * rr + 3].astype(np.float64)
                fin = np.isfinite(ref)
                assert (np.isfinite(got) == fin).all()
                # unnormalised log accumulators (Q6), magnitude ~1e3..1e5
                assert np.all(np.abs(got[fin] - ref[fin]) <= REL * np.abs(ref[fin]))
                assert np.all(np.abs(got[fin] - ref[fin]) < 2e-2)


def test_accumulators_and_mstep_match_oracle(eng):
    init, labels, utts, n_units = _ragged(cfg_seed=5, n_utt=24, T=80, L=4, n_units=4)
    corpus, model, es, om = _setup(eng, init, labels, utts, n_units)
    es.estep()
    torch.cuda.synchronize()
    stats, info = fast.estep_corpus(om, labels, utts)
    acc = es.acc.cpu().numpy().reshape(n_units, 3, om.mix, 80)
    occ = acc[..., 39]
    assert _relerr(occ, stats.occ, floor=1e-3) < REL
    assert np.allclose(acc[..., 79], occ)
    # first moments live in the standardised space of X: map the oracle's into it
    sh = es.shift.cpu().numpy()
    isc = es.inv_scale.cpu().numpy()
    sx_ref = (stats.sx - sh * stats.occ[..., None]) * isc
    assert np.all(np.abs(acc[..., :39] - sx_ref) <= REL * np.maximum(np.abs(sx_ref), np.maximum(stats.occ[..., None], 1e-2)))
    ksai, gam = es.transition_accumulators()
    fin = np.isfinite(stats.ksai_acc)
    assert (np.isfinite(ksai) == fin).all()
    assert np.all(np.abs(ksai[fin] - stats.ksai_acc[fin]) <= REL * np.abs(stats.ksai_acc[fin]))
    assert np.all(np.abs(gam - stats.gamma_acc) <= REL * np.abs(stats.gamma_acc))
    es.mstep(c_covariance=1e-6)
    torch.cuda.synchronize()
    new = fast.mstep(om, stats, c_covariance=1e-6)
    mean, var, alpha, tm = model.numpy()
    assert _relerr(alpha, new.alpha, floor=1e-3) < REL
    ok = stats.occ >= OCC_MIN
    assert np.all(np.abs(mean - new.mean)[ok] <= (REL * np.maximum(np.abs(new.mean), np.sqrt(new.var)))[ok])
    assert _var_close(var[ok], new.var[ok], utts)
    assert np.all(np.abs(tm - new.transmat) <= REL * np.maximum(new.transmat, 1e-2))


@pytest.mark.parametrize("mix", [4, 8, 16, 32, 64])
def test_accumulate_tensor_core_vs_cuda_core_and_oracle(eng, mix):
    """K3 on tcgen05 (two chained contractions, posteriors kept on chip) against the CUDA-core
    kernel and the oracle's sufficient statistics; several tiles per utterance, ragged tails."""
    n_units = 4
    truth, init, labels, utts = synth.make_corpus(16, 300, 4, n_units, mix, 50 + mix, ragged=True)
    corpus, model, es, om = _setup(eng, init, labels, utts, n_units)
    es.score()
    es.forward_backward()
    accs = {}
    for tc in (0, 1, 2):  # CUDA cores / tcgen05 with the Gaussians on the lanes and gathered blocks / tcgen05 per tile
        eng.set_option("tensor_core", min(tc, 1))
        eng.set_option("k3_kernel", 0 if tc == 2 else 1)
        es.accumulate()
        torch.cuda.synchronize()
        accs[tc] = es.acc.cpu().numpy().reshape(n_units, 3, mix, 80)
    eng.set_option("tensor_core", 1)
    eng.set_option("k3_kernel", 1)
    stats, _ = fast.estep_corpus(om, labels, utts)
    sh, isc = es.shift.cpu().numpy(), es.inv_scale.cpu().numpy()
    occ = stats.occ[..., None]
    sx_ref = (stats.sx - sh * occ) * isc
    sxx_ref = (stats.sxx - 2 * sh * stats.sx + sh * sh * occ) * isc * isc
    for tc in (0, 1, 2):
        a = accs[tc]
        assert np.all(np.abs(a[..., 39] - stats.occ) <= REL * np.maximum(stats.occ, 1e-2)), tc
        assert np.all(np.abs(a[..., 79] - stats.occ) <= REL * np.maximum(stats.occ, 1e-2)), tc
        scale = np.maximum(occ, 1e-2)  # moments of standardised features: O(occ)
        assert np.all(np.abs(a[..., :39] - sx_ref) <= REL * np.maximum(np.abs(sx_ref), scale)), tc
        assert np.all(np.abs(a[..., 40:79] - sxx_ref) <= REL * np.maximum(np.abs(sxx_ref), scale)), tc


def test_em_iteration_matches_executed_reference(eng):
    """tests/golden/estep_small.npz holds one full file-based EM iteration run by the unmodified
    reference (AcousticModel.multi_embedded_training_1/_2)."""
    g = load_golden("estep_small.npz")
    n = int(g["n_utt"])
    labels = [g[f"u{k}_label"] for k in range(n)]
    utts = [g[f"u{k}_X"] for k in range(n)]
    init = (g["mean"], g["var"], g["alpha"])
    corpus, model, es, om = _setup(eng, init, labels, utts, 3)
    es.score()
    es.forward_backward()
    torch.cuda.synchronize()
    for k in range(n):
        B = g[f"u{k}_B"]
        assert _relerr(corpus.emission_view(es.b, k).cpu().numpy(), B[1:-1]) < REL
        ab = g[f"u{k}_alpha"] + g[f"u{k}_beta"]
        lg_ref = ab - fast.lse(ab, axis=0, keepdims=True)
        got = np.exp(corpus.emission_view(es.lgam, k).cpu().numpy().astype(np.float64))
        assert np.all(np.abs(got - np.exp(lg_ref[1:-1])) <= REL * np.maximum(np.exp(lg_ref[1:-1]), 5e-2))
    es.accumulate()
    es.reduce_transitions()
    es.mstep(c_covariance=1e-6)
    torch.cuda.synchronize()
    mean, var, alpha, tm = model.numpy()
    assert _relerr(alpha, g["it1_alpha"], floor=1e-3) < REL
    stats, _ = fast.estep_corpus(om, labels, utts)
    ok = stats.occ >= OCC_MIN
    assert ok.sum() >= 0.75 * ok.size
    assert np.all(np.abs(mean - g["it1_mean"])[ok] <= (REL * np.maximum(np.abs(g["it1_mean"]), np.sqrt(g["it1_var"])))[ok])
    assert _var_close(var[ok], g["it1_var"][ok], utts)
    assert np.all(np.abs(tm - g["it1_transmat"]) <= REL * np.maximum(g["it1_transmat"], 1e-2))


def test_host_entry_point_matches_device_path(eng):
    """pc_em_iteration_host standardises like EStep.load_frames does (its own device reduction of the
    moments when no constants are passed), so both paths run the same arithmetic."""
    from poccala_b200.engine import em_iteration_host, frame_moments_host

    init, labels, utts, n_units = _ragged(cfg_seed=6, n_utt=10, T=50, L=3, n_units=4)
    utts = [x.astype(np.float32).astype(np.float64) for x in utts]  # the same fp32 frames on both paths
    corpus, model, es, om = _setup(eng, init, labels, utts, n_units)
    es.em_iteration(c_covariance=1e-6)
    torch.cuda.synchronize()
    m2, v2, a2, t2 = model.numpy()
    frames = np.concatenate(utts).astype(np.float32)
    shift, inv_scale = frame_moments_host(eng, frames)
    assert np.allclose(shift, es.shift.cpu().numpy(), rtol=1e-6, atol=1e-6)
    assert np.allclose(inv_scale, es.inv_scale.cpu().numpy(), rtol=1e-6)
    for consts in ((None, None), (shift, inv_scale)):
        mean, var, alpha = [np.ascontiguousarray(a, dtype=np.float64).copy() for a in init]
        tm = synth.default_transmat(n_units).copy()
        slp = em_iteration_host(eng, corpus, frames, mean, var, alpha, tm, c_covariance=1e-6,
                                shift=consts[0], inv_scale=consts[1])
        assert abs(slp - float(es.utt_logp.sum())) < 1e-6 * abs(slp)
        assert np.allclose(mean, m2, rtol=1e-5, atol=1e-6)
        assert np.allclose(var, v2, rtol=1e-4, atol=1e-7)
        assert np.allclose(alpha, a2, rtol=1e-5)
        assert np.allclose(tm, t2, rtol=1e-5, atol=1e-9)


def _offset_problem(cfg_seed, n_utt, T, L, n_units, mix=4):
    """Features far from the origin and on very different scales (dimension 0 sits at +50, every
    third dimension is stretched x12, |x| < 100 for the reference's +100 bias, Q7): the expanded
    quadratic of the scoring contraction only survives this through the standardisation."""
    truth, init, labels, utts = synth.make_corpus(n_utt, T, L, n_units, mix, cfg_seed, ragged=True)
    D = init[0].shape[-1]
    scale = np.ones(D)
    scale[2::3] = 8.0
    offset = np.zeros(D)
    offset[0], offset[5] = 50.0, -35.0
    utts = [x * scale + offset for x in utts]
    init = (init[0] * scale + offset, init[1] * scale * scale, init[2])
    assert max(np.abs(x).max() for x in utts) < 100.0, "Q7: the reference adds 100 before taking logs"
    return init, labels, utts


def test_host_entry_point_offset_data_matches_oracle(eng):
    """The end-to-end entry point (the call bench.py times as `e2e`) against the fp64 oracle on data
    with a large mean offset: log-likelihoods and re-estimated parameters within 1e-4."""
    from poccala_b200.engine import em_iteration_host, frame_moments_host

    n_units = 4
    init, labels, utts = _offset_problem(31, 30, 80, 3, n_units)
    corpus = _corpus(eng, labels, utts, n_units)
    tm0 = synth.default_transmat(n_units)
    om = fast.Model(*init, tm0)
    stats, info = fast.estep_corpus(om, labels, utts)
    new = fast.mstep(om, stats, c_covariance=1e-6)
    frames = np.concatenate(utts).astype(np.float32)
    shift, inv_scale = frame_moments_host(eng, frames)
    for consts in ((None, None), (shift, inv_scale)):
        mean, var, alpha = [np.ascontiguousarray(a, dtype=np.float64).copy() for a in init]
        tm = tm0.copy()
        slp = em_iteration_host(eng, corpus, frames, mean, var, alpha, tm, c_covariance=1e-6,
                                shift=consts[0], inv_scale=consts[1])
        assert abs(slp - float(np.sum(info["logp"]))) <= 1e-5 * abs(slp), (slp, float(np.sum(info["logp"])))
        ok = stats.occ >= OCC_MIN
        assert ok.sum() >= 0.6 * ok.size, ok.mean()
        assert _relerr(alpha, new.alpha, floor=1e-3) < REL, _relerr(alpha, new.alpha, floor=1e-3)
        sd = np.sqrt(new.var)
        # the stated bound (relative to the parameter, as in the other tests) ...
        merr = (np.abs(mean - new.mean) / np.maximum(np.abs(new.mean), sd))[ok].max()
        assert merr <= REL, merr
        # ... and what the standardised contraction actually reaches: the error does not grow with the offset
        # (measured 2e-4 of the distance to the corpus centre; without standardisation it is ~1e-1)
        merr_c = (np.abs(mean - new.mean) / np.maximum(np.abs(new.mean - shift), sd))[ok].max()
        assert merr_c <= 5e-4, merr_c
        gvar = np.concatenate(utts, axis=0).var(axis=0)
        # A variance is a posterior-weighted average of (x - mu)^2.  The fp32 posteriors carry a relative error of up
        # to 1e-4 each (the stated bound for posteriors: log gamma is an fp32 difference of numbers of magnitude 1e3);
        # in the average that error shrinks with the square root of the number of frames that share the weight, and
        # ((x - mu)^2 - var) spreads by sqrt(2) var.  So the 1e-4 bound holds from four frames of occupancy on, and
        # below that it is 1e-4 * 2 / sqrt(occupancy) (measured: 1.2e-4 at 1.4 frames, 1.3e-4 at 0.14 frames,
        # 2.4e-4 at 3e-3 frames), down to a hundredth of a frame.
        vrel = np.abs(var - new.var) / np.maximum(new.var, 1e-2 * gvar)
        ok_v = stats.occ >= 1e-2
        assert (stats.occ >= 4.0).sum() >= 0.3 * ok_v.size
        vtol = REL * np.maximum(1.0, 2.0 / np.sqrt(np.maximum(stats.occ, 1e-2)))[..., None]
        worst = np.unravel_index(np.argmax(np.where(ok_v[..., None], vrel / vtol, 0.0)), vrel.shape)
        assert np.all((vrel <= vtol)[ok_v]), (vrel[worst], worst, stats.occ[worst[:3]], new.var[worst], var[worst])
        terr = (np.abs(tm - new.transmat) / np.maximum(new.transmat, 1e-2)).max()
        assert terr <= REL, terr


def test_host_entry_point_reports_unstandardised_frames(eng):
    """Constants that do not describe the frames push standardised values past +-240: the call fails
    instead of saturating silently."""
    from poccala_b200._native import NativeError
    from poccala_b200.engine import em_iteration_host

    n_units = 4
    init, labels, utts = _offset_problem(32, 6, 40, 2, n_units)
    corpus = _corpus(eng, labels, utts, n_units)
    mean, var, alpha = [np.ascontiguousarray(a, dtype=np.float64).copy() for a in init]
    tm = synth.default_transmat(n_units).copy()
    D = mean.shape[-1]
    with pytest.raises(NativeError, match="clamped"):
        em_iteration_host(eng, corpus, np.concatenate(utts).astype(np.float32), mean, var, alpha, tm,
                          c_covariance=1e-6, shift=np.zeros(D), inv_scale=np.full(D, 10.0))
    assert eng.get_option("clamped") == 0  # reported once


def _viterbi_check(eng, labels, utts, init, n_units, transmat=None, emissions64=None):
    from poccala_b200.engine import host_log_bands, viterbi

    corpus, model, es, om = _setup(eng, init, labels, utts, n_units, transmat=transmat)
    ls, ln = host_log_bands(om.transmat, eng.device)
    logpi = torch.as_tensor(np.array([np.log(np.ones(3 * len(l) + 2) / (3 * len(l) + 2))[0] for l in labels])).to(eng.device)
    if emissions64 is None:
        es.score()
        b = es.b
    else:
        b = emissions64
    score, path, units = viterbi(eng, corpus, b, ls, ln, utt_logpi=logpi)
    torch.cuda.synchronize()
    score, path, units = score.cpu().numpy(), path.cpu().numpy(), units.cpu().numpy()
    for u, lab in enumerate(labels):
        e = corpus.emission_view(b, u).cpu().numpy().astype(np.float64)  # identical emission scores
        ols, oln = fast.banded_transitions(om, np.asarray(lab)[None])
        sc, pa = fast.viterbi_banded(ols, oln, fast.full_emissions(e.T[None]))
        f0, f1 = corpus.frame_off[u], corpus.frame_off[u + 1]
        assert score[u] == sc[0], (u, score[u], sc[0])  # bit-exact fp64
        assert (path[f0:f1] == pa[0]).all()
        unit_of_state = np.array([lab[0]] + [x for x in lab for _ in range(3)] + [lab[-1]])
        assert (units[f0:f1] == unit_of_state[pa[0]]).all()
    return corpus


def test_viterbi_bit_exact_ragged(eng):
    init, labels, utts, n_units = _ragged(cfg_seed=7, n_utt=30, T=120, L=6, n_units=6)
    _viterbi_check(eng, labels, utts, init, n_units)


def test_viterbi_bit_exact_two_states_per_lane(eng):
    truth, init, labels, utts = synth.make_corpus(6, 200, 20, 8, 4, 9)  # N = 62 > 32 lanes
    _viterbi_check(eng, labels, utts, init, 8)


def test_viterbi_single_frame_and_trained_transitions(eng):
    truth, init, labels, utts = synth.make_corpus(5, 40, 3, 4, 4, 10, ragged=True)
    utts[0] = utts[0][:1]  # T = 1
    rng = np.random.default_rng(0)
    tm = synth.default_transmat(4)
    for u in range(4):
        for j in range(1, 4):
            a = rng.uniform(0.2, 0.8)
            tm[u, j, j], tm[u, j, j + 1] = a, 1 - a
    _viterbi_check(eng, labels, utts, init, 4, transmat=tm)


def test_viterbi_ties_golden(eng):
    """Integer emissions force exact ties; expected paths come from the executed reference."""
    from poccala_b200.engine import Corpus, host_log_bands, viterbi

    v = load_golden("viterbi_ties.npz")
    for k in range(int(v["n"])):
        A, B = v[f"v{k}_A"], v[f"v{k}_B"]
        N, T = B.shape
        L = (N - 2) // 3
        corpus = Corpus(eng, [np.zeros(L, dtype=np.int32)], np.array([T], dtype=np.int32), 1)
        sp = (3 * L + 7) & ~7
        buf = np.zeros((T, sp))
        buf[:, :3 * L] = B[1:-1].T
        b64 = torch.as_tensor(buf.reshape(-1)).to(eng.device)
        tm = np.zeros((1, 5, 5))
        tm[0, 0, 1] = 1.0
        for j in range(1, 4):
            tm[0, j, j] = tm[0, j, j + 1] = 0.5
        ls, ln = host_log_bands(tm, eng.device)
        logpi = torch.as_tensor(np.array([np.log(np.ones(N) / N)[0]])).to(eng.device)
        score, path, _ = viterbi(eng, corpus, b64, ls, ln, utt_logpi=logpi)
        assert float(score[0]) == float(v[f"v{k}_score"])
        assert (path.cpu().numpy() == v[f"v{k}_path"].astype(np.int64)).all()


def test_dense_scoring_matches_oracle(eng):
    rng = np.random.default_rng(11)
    n_states, mix, F = 7, 8, 300
    mean = rng.normal(size=(n_states * mix, 39))
    var = rng.uniform(0.5, 1.5, size=(n_states * mix, 39))
    alpha = rng.dirichlet(np.full(mix, 2.0), size=n_states).reshape(-1)
    x = rng.normal(size=(F, 39))
    dev = eng.device
    W = eng.pack_gmm(torch.as_tensor(mean).to(dev), torch.as_tensor(var).to(dev), torch.as_tensor(alpha).to(dev))
    X = eng.prepare_rows(torch.as_tensor(x).to(dev))
    out = eng.score_dense(X, W, n_states, mix).cpu().numpy()
    d = x[:, None, :] - mean[None]
    c = np.log(alpha) - 39 / 2 * np.log(2 * np.pi) - 0.5 * var.sum(-1) - 0.5 * (d * d / var).sum(-1)
    ref = fast.lse(c.reshape(F, n_states, mix), axis=-1)
    assert _relerr(out, ref) < REL


@pytest.mark.parametrize("mix,n_states", [(16, 40), (64, 7), (4, 3)])
def test_dense_scoring_tensor_core_matches_oracle(eng, mix, n_states):
    """configs[2] shape at test size: every frame against every state of one model through the tcgen05
    kernel (frames in groups of 384 labelled with all pseudo units), against the fp64 closed form and
    the CUDA-core dense kernel."""
    rng = np.random.default_rng(100 + mix)
    F = 1000
    mean = rng.normal(size=(n_states * mix, 39))
    var = rng.uniform(0.5, 1.5, size=(n_states * mix, 39))
    alpha = rng.dirichlet(np.full(mix, 2.0), size=n_states).reshape(-1)
    x = rng.normal(size=(F, 39))
    dev = eng.device
    t = lambda a: torch.as_tensor(a).to(dev)
    out = eng.score_dense_tc(t(x), t(mean), t(var), t(alpha), mix).cpu().numpy()
    d = x[:, None, :] - mean[None]
    c = np.log(alpha) - 39 / 2 * np.log(2 * np.pi) - 0.5 * var.sum(-1) - 0.5 * (d * d / var).sum(-1)
    ref = fast.lse(c.reshape(F, n_states, mix), axis=-1)
    assert out.shape == ref.shape
    assert _relerr(out, ref) < REL
    W = eng.pack_gmm(t(mean), t(var), t(alpha))
    simt = eng.score_dense(eng.prepare_rows(t(x)), W, n_states, mix).cpu().numpy()
    assert np.abs(out - simt).max() < 2e-3


def test_error_behaviour(eng):
    from poccala_b200 import _native as nat
    from poccala_b200.engine import Corpus

    with pytest.raises(ValueError):
        eng.prepare_rows(torch.zeros((4, 40), device=eng.device))  # DataDimensionError analogue
    with pytest.raises(nat.NativeError):
        Corpus(eng, [np.array([0, 7])], np.array([10], dtype=np.int32), 3)  # label outside the unit set
    with pytest.raises(nat.NativeError):
        Corpus(eng, [np.array([0])], np.array([0], dtype=np.int32), 3)  # empty utterance


@pytest.mark.parametrize("L", [3, 14])  # one state per lane / two states per lane in K2
def test_activity_flags_from_forward_backward_equal_the_pre_pass(eng, L):
    """K2's log-gamma helper sets K3's (tile, position) activity flags while it writes the rows; K3's
    own pre-pass derives them from log gamma that came from elsewhere.  Same flags and the same
    statistics either way, including utterances of 3, 127, 128, 129 and 300 frames."""
    from poccala_b200 import _native as nat

    n_units, mix = 5, 4
    truth, init, _, _ = synth.make_corpus(2, 8, L, n_units, mix, 7)
    rng = np.random.default_rng(70 + L)
    lens = [3, 5, 127, 128, 129, 255, 256, 257, 300, 40, 77]
    labels = [rng.integers(0, n_units, size=int(rng.integers(1, L + 1))).astype(np.int32) for _ in lens]
    labels[0] = labels[0][:1]
    utts = [synth.make_utterance(lab, T, truth, 900 + i) for i, (lab, T) in enumerate(zip(labels, lens))]
    corpus, model, es, om = _setup(eng, init, labels, utts, n_units)
    es.score()
    es.forward_backward()
    es.accumulate()  # flags set by K2
    torch.cuda.synchronize()
    n_k2 = int(nat.lib().pc_corpus_active_tiles(corpus.c))
    acc_k2 = es.acc.clone()
    es.accumulate()  # the flags were consumed: K3's own pre-pass
    torch.cuda.synchronize()
    n_pre = int(nat.lib().pc_corpus_active_tiles(corpus.c))
    assert n_k2 == n_pre and 0 < n_k2 <= int(nat.lib().pc_corpus_total_tiles(corpus.c))
    scale = es.acc.abs().amax(dim=1, keepdim=True).clamp_min(1e-6)
    assert ((acc_k2 - es.acc).abs() <= 1e-9 * scale).all()  # fp64 atomics: order only
    stats, info = fast.estep_corpus(om, labels, utts)
    occ = acc_k2.cpu().numpy().reshape(n_units, 3, om.mix, 80)[..., 39]
    assert _relerr(occ, stats.occ, floor=1e-3) < REL
