"""CPU oracle — vectorised fp64 restatement for sizes where the
loop-faithful port (oracle/ref_port.py) would need core-hours.

TEST INFRASTRUCTURE ONLY — same rules as oracle/ref_port.py: never imported by the product.
Pinned by ``tests/test_oracle_golden.py`` to the executed reference (tests/golden/*.npz) and to
ref_port on seeded inputs.

Differences from ref_port are purely computational (banded instead of dense recursions, linear
instead of log-domain GMM statistics, batches of equal-shape utterances); the arithmetic being
restated is cited per function (paths relative to /root/reference).
"""
from __future__ import annotations

import math

import numpy as np

np.seterr(all="ignore")

LOG_2PI = np.log(2 * np.pi)
E = 3
S = 5
BW_THRESHOLD = 0.64
BIAS = 100.0


class Model:
    """Dense parameter arrays: mean/var [U,3,M,D], alpha [U,3,M], transmat [U,5,5]."""

    def __init__(self, mean, var, alpha, transmat):
        self.mean = np.asarray(mean, dtype=np.float64)
        self.var = np.asarray(var, dtype=np.float64)
        self.alpha = np.asarray(alpha, dtype=np.float64)
        self.transmat = np.asarray(transmat, dtype=np.float64)
        self.n_units, _, self.mix, self.dim = self.mean.shape

    def copy(self):
        return Model(self.mean.copy(), self.var.copy(), self.alpha.copy(), self.transmat.copy())

    def packed(self):
        """A.1 contraction form: c = x.(mu/var) + x^2.(-1/(2 var)) + k,
        k = log alpha - D/2 log 2pi - 1/2 sum var (Q1, util.py:29) - 1/2 sum mu^2/var."""
        w1 = self.mean / self.var
        w2 = -0.5 / self.var
        k = (np.log(self.alpha) - self.dim / 2 * LOG_2PI - 0.5 * self.var.sum(-1)
             - 0.5 * (self.mean ** 2 / self.var).sum(-1))
        return w1, w2, k


def lse(a, axis=-1, keepdims=False):
    """util.py:54-77 semantics, vectorised: +-inf maxima are returned unchanged."""
    mx = np.max(a, axis=axis, keepdims=True)
    safe = np.where(np.isfinite(mx), mx, 0.0)
    out = safe + np.log(np.sum(np.exp(a - safe), axis=axis, keepdims=True))
    out = np.where(np.isfinite(mx), out, mx)
    return out if keepdims else np.squeeze(out, axis=axis)


def logaddexp(a, b):
    mx = np.maximum(a, b)
    safe = np.where(np.isfinite(mx), mx, 0.0)
    out = safe + np.log(np.exp(a - safe) + np.exp(b - safe))
    return np.where(np.isfinite(mx), out, mx)


# --------------------------------------------------------------------------- scoring (A.1)
def score_components(model, labels, X):
    """labels [B,L] unit ids, X [B,T,D] -> c [B,T,3L,M] component log-scores (log alpha + log N)."""
    w1, w2, k = model.packed()
    B, T, D = X.shape
    L = labels.shape[1]
    w1g = w1[labels].reshape(B, E * L * model.mix, D)
    w2g = w2[labels].reshape(B, E * L * model.mix, D)
    kg = k[labels].reshape(B, 1, E * L * model.mix)
    c = np.matmul(X, w1g.transpose(0, 2, 1)) + np.matmul(X * X, w2g.transpose(0, 2, 1)) + kg
    return c.reshape(B, T, E * L, model.mix)


def score_components_direct(model, labels, X):
    """Same as score_components but in the reference's direct form -1/2 sum (x-mu)^2/var."""
    B, T, D = X.shape
    L = labels.shape[1]
    mu = model.mean[labels].reshape(B, 1, E * L, model.mix, D)
    var = model.var[labels].reshape(B, 1, E * L, model.mix, D)
    al = model.alpha[labels].reshape(B, 1, E * L, model.mix)
    d = X[:, :, None, None, :] - mu
    return (np.log(al) - D / 2 * LOG_2PI - 0.5 * var.sum(-1)) + (-0.5 * (d * (1.0 / var) * d).sum(-1))


# --------------------------------------------------------------------------- sentence HMM (A.2)
def banded_transitions(model, labels):
    """AcousticModel.py:978-988 as two bands: log a_self[B,N], log a_next[B,N] (i -> i+1).
    Entry state 0: self = log 0, next = log transmat_{u0}[0,1]; exit state N-1: both log 0."""
    B, L = labels.shape
    N = E * L + 2
    tm = model.transmat[labels]  # [B,L,5,5]
    a_self = np.zeros((B, N))
    a_next = np.zeros((B, N))
    a_next[:, 0] = tm[:, 0, 0, 1]
    a_self[:, 0] = tm[:, 0, 0, 0]
    for r in range(E):
        a_self[:, 1 + r : 1 + E * L : E] = tm[:, :, 1 + r, 1 + r]
        a_next[:, 1 + r : 1 + E * L : E] = tm[:, :, 1 + r, 2 + r]
    return np.log(a_self), np.log(a_next)


def full_emissions(b_emit):
    """b_emit [B,T,3L] -> Bfull [B,T,N] with virtual rows 0 (log 1) and N-1 (log 0) (Q3)."""
    B, T, n = b_emit.shape
    out = np.empty((B, T, n + 2))
    out[:, :, 0] = 0.0
    out[:, :, 1:-1] = b_emit
    out[:, :, -1] = -np.inf
    return out


def backward_banded(ls, ln, Bf):
    """LHMM.py:353-366 restricted to the two non-zero bands."""
    B, T, N = Bf.shape
    beta = np.zeros((B, T, N))
    for t in range(T - 2, -1, -1):
        nb = Bf[:, t + 1] + beta[:, t + 1]
        stay = ls + nb
        move = np.full((B, N), -np.inf)
        move[:, :-1] = ln[:, :-1] + nb[:, 1:]
        beta[:, t] = logaddexp(stay, move)
    return beta


def forward_banded(ls, ln, Bf, logpi):
    """LHMM.py:335-351 restricted to the two non-zero bands."""
    B, T, N = Bf.shape
    alpha = np.empty((B, T, N))
    alpha[:, 0] = logpi + Bf[:, 0]
    for t in range(1, T):
        prev = alpha[:, t - 1]
        stay = prev + ls
        move = np.full((B, N), -np.inf)
        move[:, 1:] = prev[:, :-1] + ln[:, :-1]
        alpha[:, t] = logaddexp(stay, move) + Bf[:, t]
    return alpha


def baum_welch_batch(ls, ln, Bf, max_iter=100):
    """LHMM.py:526-544 for a batch: per utterance iterate pi := exp(alpha_0+beta_0-lse) while the
    forward log-likelihood grows by more than 0.64 (Q5).  Returns alpha (of the last executed
    iteration per utterance), beta, logp, iters, logpi_used."""
    B, T, N = Bf.shape
    beta = backward_banded(ls, ln, Bf)
    logpi = np.log(np.ones((B, N)) / N)
    q = np.full(B, -np.inf)
    active = np.ones(B, dtype=bool)
    alpha = np.empty((B, T, N))
    logp = np.empty(B)
    iters = np.zeros(B, dtype=np.int64)
    used = logpi.copy()
    for _ in range(max_iter):
        idx = np.where(active)[0]
        if len(idx) == 0:
            break
        a = forward_banded(ls[idx], ln[idx], Bf[idx], logpi[idx])
        qn = lse(a[:, -1], axis=-1)
        alpha[idx] = a
        logp[idx] = qn
        used[idx] = logpi[idx]
        iters[idx] += 1
        g0 = a[:, 0] + beta[idx, 0]
        newpi = np.exp(g0 - lse(g0, axis=-1, keepdims=True))  # linear space: underflows to exact 0
        logpi[idx] = np.log(newpi)
        grow = (qn - q[idx]) > BW_THRESHOLD
        q[idx] = np.where(grow, qn, q[idx])
        active[idx] = grow
    return alpha, beta, logp, iters, used


def transition_stats(ls, ln, Bf, alpha, beta):
    """LHMM.py:431-445: unnormalised log ksai on the two bands and log gamma, sums over t < T-1."""
    nb = Bf[:, 1:] + beta[:, 1:]  # [B,T-1,N]
    a = alpha[:, :-1]
    k_self = lse(a + ls[:, None] + nb, axis=1)
    mv = np.full_like(a, -np.inf)
    mv[:, :, :-1] = a[:, :, :-1] + ln[:, None, :-1] + nb[:, :, 1:]
    k_next = lse(mv, axis=1)
    gam = lse(a + beta[:, :-1], axis=1)
    if a.shape[1] == 0:
        k_self = np.full(ls.shape, -np.inf)
        k_next = np.full(ls.shape, -np.inf)
        gam = np.full(ls.shape, -np.inf)
    return k_self, k_next, gam


class Stats:
    """EM sufficient statistics.  GMM part linear (A.4), transition part log-domain (Q6)."""

    def __init__(self, model):
        U, M, D = model.n_units, model.mix, model.dim
        self.occ = np.zeros((U, E, M))
        self.sx = np.zeros((U, E, M, D))
        self.sxx = np.zeros((U, E, M, D))
        self.socc = np.zeros((U, E))
        self.ksai_acc = np.full((U, E, S), -np.inf)
        self.gamma_acc = np.full((U, E), -np.inf)
        self.logp = 0.0
        self.frames = 0
        self._pending_k = []
        self._pending_g = []
        self._pending_u = []

    def add_transitions(self, labels, k_self, k_next, gam):
        B, L = labels.shape
        ks = np.full((B, L, E, S), -np.inf)
        for r in range(E):
            ks[:, :, r, r + 1] = k_self[:, 1 + r : 1 + E * L : E]
            ks[:, :, r, r + 2] = k_next[:, 1 + r : 1 + E * L : E]
        g = np.stack([gam[:, 1 + r : 1 + E * L : E] for r in range(E)], axis=-1)
        self._pending_k.append(ks.reshape(B * L, E, S))
        self._pending_g.append(g.reshape(B * L, E))
        self._pending_u.append(labels.reshape(-1))

    def finalize(self):
        if not self._pending_u:
            return
        ks = np.concatenate(self._pending_k)
        g = np.concatenate(self._pending_g)
        u = np.concatenate(self._pending_u)
        for unit in np.unique(u):
            sel = u == unit
            self.ksai_acc[unit] = logaddexp(self.ksai_acc[unit], lse(ks[sel], axis=0))
            self.gamma_acc[unit] = logaddexp(self.gamma_acc[unit], lse(g[sel], axis=0))
        self._pending_k, self._pending_g, self._pending_u = [], [], []

    def merge(self, other):
        """What the cross-rank reduction does: sums for linear parts, log-add for log parts."""
        self.finalize()
        other.finalize()
        self.occ += other.occ
        self.sx += other.sx
        self.sxx += other.sxx
        self.socc += other.socc
        self.ksai_acc = logaddexp(self.ksai_acc, other.ksai_acc)
        self.gamma_acc = logaddexp(self.gamma_acc, other.gamma_acc)
        self.logp += other.logp
        self.frames += other.frames

    def log_domain(self, model):
        """The reference's own accumulator values (Clustering.py:653-680) implied by the linear
        statistics: acc, alpha_acc, mean_acc (+100 bias), covariance_acc (around the OLD mean)."""
        self.finalize()
        mu = model.mean
        return dict(
            acc=np.log(self.occ),
            alpha_acc=np.log(self.socc),
            mean_acc=np.log(self.sx + BIAS * self.occ[..., None]),
            cov_acc=np.log(self.sxx - 2 * mu * self.sx + mu * mu * self.occ[..., None]),
        )


def estep_batch(model, labels, X, stats=None, keep=False, direct=False):
    """A.3 + A.4 for a batch of equal-shape utterances: labels [B,L], X [B,T,D]."""
    labels = np.asarray(labels)
    X = np.asarray(X, dtype=np.float64)
    B, T, D = X.shape
    L = labels.shape[1]
    c = (score_components_direct if direct else score_components)(model, labels, X)
    b = lse(c, axis=-1)  # [B,T,3L]
    Bf = full_emissions(b)
    ls, ln = banded_transitions(model, labels)
    alpha, beta, logp, iters, logpi_used = baum_welch_batch(ls, ln, Bf)
    ab = alpha + beta
    lgam_full = ab - lse(ab, axis=-1, keepdims=True)  # LHMM.py:486-500, all t
    lgam = lgam_full[:, :, 1:-1]
    k_self, k_next, gam = transition_stats(ls, ln, Bf, alpha, beta)
    if stats is not None:
        post = np.exp(lgam[..., None] + c - b[..., None])  # [B,T,3L,M]
        M = model.mix
        P = post.reshape(B, T, E * L * M)
        occ = P.sum(1).reshape(B, L, E, M)
        sx = np.matmul(P.transpose(0, 2, 1), X).reshape(B, L, E, M, D)
        sxx = np.matmul(P.transpose(0, 2, 1), X * X).reshape(B, L, E, M, D)
        so = np.exp(lgam).sum(1).reshape(B, L, E)
        np.add.at(stats.occ, labels, occ)
        np.add.at(stats.sx, labels, sx)
        np.add.at(stats.sxx, labels, sxx)
        np.add.at(stats.socc, labels, so)
        stats.add_transitions(labels, k_self, k_next, gam)
        stats.logp += float(logp.sum())
        stats.frames += B * T
    out = dict(b=b, lgam=lgam, logp=logp, iters=iters, k_self=k_self, k_next=k_next, gamma=gam)
    if keep:
        out.update(alpha=alpha, beta=beta, c=c, logpi_used=logpi_used, ls=ls, ln=ln, Bf=Bf)
    return out


def estep_corpus(model, labels_list, X_list, chunk=64, direct=False):
    """Ragged corpus: group by (T, L), run estep_batch per group, return (Stats, per-utt dict)."""
    stats = Stats(model)
    n = len(X_list)
    logp = np.empty(n)
    iters = np.empty(n, dtype=np.int64)
    groups = {}
    for i, (lab, X) in enumerate(zip(labels_list, X_list)):
        groups.setdefault((len(X), len(lab)), []).append(i)
    per_utt = [None] * n
    for (_, _), ids in groups.items():
        for lo in range(0, len(ids), chunk):
            sub = ids[lo : lo + chunk]
            r = estep_batch(model, np.stack([labels_list[i] for i in sub]), np.stack([X_list[i] for i in sub]),
                            stats, direct=direct)
            for k, i in enumerate(sub):
                logp[i] = r["logp"][k]
                iters[i] = r["iters"][k]
                per_utt[i] = dict(b=r["b"][k], lgam=r["lgam"][k])
    stats.finalize()
    return stats, dict(logp=logp, iters=iters, per_utt=per_utt)


def mstep(model, stats, c_covariance=1e-3, fix_code=0):
    """A.5: LHMM.py:509-524 + Clustering.py:682-693 from linear statistics.  Units that never
    occurred keep -inf accumulators in the reference and are not re-estimated by this oracle."""
    stats.finalize()
    new = model.copy()
    seen = np.isfinite(stats.gamma_acc).any(axis=1) | (stats.socc.sum(axis=1) > 0)
    if not (fix_code & 4):
        tm = np.exp(stats.ksai_acc - stats.gamma_acc[..., None])
        new.transmat[seen, 1:-1, :] = tm[seen]
    if not (fix_code & 2):
        occ = stats.occ[..., None]
        alpha = stats.occ / stats.socc[..., None]
        mean = stats.sx / occ
        var = (stats.sxx - 2 * model.mean * stats.sx + model.mean ** 2 * occ) / occ
        var = np.where(var < c_covariance, c_covariance, var)
        new.alpha[seen] = alpha[seen]
        new.mean[seen] = mean[seen]
        new.var[seen] = var[seen]
    return new


# --------------------------------------------------------------------------- Viterbi (A.6)
def viterbi_banded(ls, ln, Bf):
    """LHMM.py:546-609 on the two bands, fp64, batch [B,T,N].  Ties -> lower index (j-1); cells
    with both candidates -inf store backpointer 0 like the reference's first-argmax of an all -inf
    column.  Returns (score [B], path [B,T] int64)."""
    B, T, N = Bf.shape
    p = np.log(np.ones((B, N)) / N) + Bf[:, 0]
    bp = np.zeros((B, T, N), dtype=np.int64)
    j = np.arange(N)[None, :]
    for t in range(1, T):
        stay = p + ls
        move = np.full((B, N), -np.inf)
        move[:, 1:] = p[:, :-1] + ln[:, :-1]
        take_move = move >= stay
        best = np.where(take_move, move, stay)
        idx = np.where(take_move, j - 1, j)
        idx = np.where(np.isneginf(best), 0, idx)
        idx = np.where(np.isnan(best), 0, idx)
        bp[:, t] = idx
        p = best + Bf[:, t]
    end = np.argmax(p, axis=1)  # first maximum
    score = p[np.arange(B), end]
    path = np.zeros((B, T), dtype=np.int64)
    cur = end.copy()
    for t in range(T - 1, -1, -1):
        path[:, t] = cur
        cur = bp[np.arange(B), t, cur]
    return score, path


def viterbi_corpus(model, labels_list, X_list, chunk=64, emissions=None):
    """Forced alignment of a ragged corpus.  ``emissions`` (list of [T,3L] arrays) overrides the
    oracle's own scoring so that paths can be compared given IDENTICAL emission scores."""
    n = len(X_list)
    scores = np.empty(n)
    paths = [None] * n
    groups = {}
    for i, (lab, X) in enumerate(zip(labels_list, X_list)):
        groups.setdefault((len(X), len(lab)), []).append(i)
    for _, ids in groups.items():
        for lo in range(0, len(ids), chunk):
            sub = ids[lo : lo + chunk]
            labels = np.stack([labels_list[i] for i in sub])
            if emissions is None:
                X = np.stack([X_list[i] for i in sub]).astype(np.float64)
                b = lse(score_components(model, labels, X), axis=-1)
            else:
                b = np.stack([np.asarray(emissions[i], dtype=np.float64) for i in sub])
            ls, ln = banded_transitions(model, labels)
            sc, pa = viterbi_banded(ls, ln, full_emissions(b))
            for k, i in enumerate(sub):
                scores[i] = sc[k]
                paths[i] = pa[k]
    return scores, paths


# --------------------------------------------------------------------------- k-means (A.7)
def kmeans_compat(data, k, rng_random, max_passes=10 ** 9):
    """A.7 with a vectorised metric (dimension 0 only, Q2).  ``rng_random`` is a ``random.Random``
    (or the ``random`` module) already seeded.  Returns dict(mean, var, alpha, owner, members)
    where members[k] lists point indices in the cluster's insertion order (seed first)."""
    X = np.asarray(data, dtype=np.float64)
    n = len(X)
    x0 = X[:, 0]
    owner = np.full(n, -1, dtype=np.int64)
    keyed = np.zeros(n, dtype=bool)  # True once the point holds a non-seed key
    seeds = []
    c0 = rng_random.randint(0, n - 1)
    seeds.append(c0)
    owner[c0] = 0
    dl = (np.abs(x0[c0] - x0) ** 2) ** 0.5
    total = 0.0
    for v in dl:
        total += v
    if total == 0.0:
        idx = rng_random.sample(range(0, n), k - 1)
        assert k <= 2, "reference asserts here for k > 2 (Clustering.py:1004-1008)"
        for kk in range(1, k):
            for i in idx:
                seeds.append(i)
                owner[i] = kk
    else:
        for kk in range(1, k):
            r = rng_random.randint(0, int(total))
            rr = r  # sequential subtraction exactly like Clustering.py:1013-1018
            pick = -1
            for i in range(n):
                rr -= dl[i]
                if rr < 0:
                    pick = i
                    break
            assert pick >= 0
            seeds.append(pick)
            owner[pick] = kk
    members = [[s] for s in seeds]
    centre = np.array([X[s].copy() for s in seeds])
    moved = True
    passes = 0
    while moved and passes < max_passes:
        moved = False
        passes += 1
        for kk in range(k):
            dist = (np.abs(centre[kk, 0] - x0) ** 2) ** 0.5
            own = np.where(owner >= 0, (np.abs(centre[np.maximum(owner, 0), 0] - x0) ** 2) ** 0.5, np.inf)
            elig = (owner != kk) & ((owner < 0) | (own > dist))
            if not elig.any():
                break
            cand = np.where(elig, dist, np.inf)
            bi = int(np.argmin(cand))  # first strictly smallest
            moved = True
            if owner[bi] != -1 and keyed[bi]:
                old = members[owner[bi]]  # drop the keyed (latest) entry, never the seed entry
                del old[len(old) - 1 - old[::-1].index(bi)]
            rng_random.random()  # the dict key draw (Clustering.py:932); collisions are ~1e-15 events
            keyed[bi] = True
            owner[bi] = kk
            members[kk].append(bi)
        for kk in range(k):
            c = np.zeros(X.shape[1])
            for i in members[kk]:
                c += X[i]
            centre[kk] = c / len(members[kk])
    var = np.empty_like(centre)
    for kk in range(k):
        pts = X[members[kk]]
        v = np.zeros(X.shape[1])
        for p in pts:
            v += (centre[kk] - p) ** 2
        v /= len(pts)
        v = np.where(v < 1e-4, 1e-4, v)
        var[kk] = (v ** 0.5) ** 2
    alpha = np.array([len(m) / n for m in members])
    return dict(mean=centre, var=var, alpha=alpha, owner=owner, members=members, passes=passes, seeds=seeds)


# --------------------------------------------------------------------------- alignment post-processing (§8 f3)
def split_segment(n, parts=E):
    """AcousticModel.__eq_segment(mode='g') (AcousticModel.py:613-626): a segment of n frames cut
    into `parts` slices of n // parts frames, the last slice taking the remainder.  Returns the
    (start, end) offsets of the slices (empty slices when n < parts)."""
    chunk = n // parts
    cuts = [(i * chunk, (i + 1) * chunk) for i in range(parts - 1)]
    cuts.append(((parts - 1) * chunk, n))
    return cuts


def segment_uniform(label, T):
    """AcousticModel.__eq_segment(mode='e') (AcousticModel.py:605-612): T frames cut into len(label)
    chunks of T // len(label) frames in label order; the remainder is never saved.
    Returns [(unit, start, end)]."""
    chunk = T // len(label)
    return [(u, p * chunk, (p + 1) * chunk) for p, u in enumerate(label)]


def discriminate(unit, sequence):
    """AcousticModel.discriminate (AcousticModel.py:937-955): the frames labelled `unit`, one index
    array per maximal run of consecutive frames (time order)."""
    loc = np.flatnonzero(np.asarray(sequence) == unit)
    runs = {}
    for rank, f in enumerate(loc):
        runs.setdefault(int(f) - rank, []).append(int(f))  # consecutive frames share f - rank
    return [np.array(runs[k]) for k in sorted(runs)]


def segment_alignment(label, unit_sequence):
    """multi_process_data after Viterbi (AcousticModel.py:750-764): None when the aligned sequence
    holds fewer distinct units than the label (the utterance is discarded, :753-757), else
    [(unit, start, end)] for every run of every unit, in time order."""
    seq = np.asarray(unit_sequence)
    if len(set(seq.tolist())) < len(set(label)):
        return None
    out = []
    for u in set(label):
        for run in discriminate(u, seq):
            out.append((u, int(run[0]), int(run[-1]) + 1))
    return sorted(out, key=lambda s: s[1])


def state_frames(segments_per_utt, frame_off, n_units):
    """AcousticModel.__get_gmmdata (AcousticModel.py:630-644) over a corpus: for every (unit, state)
    the corpus frame indices of its data set - each unit segment cut by `split_segment`, slices
    concatenated in (utterance, time) order.  segments_per_utt[u] = [(unit, start, end)] or None.
    Returns a list indexed by unit * 3 + state of int64 arrays."""
    sets = [[] for _ in range(n_units * E)]
    for u, segs in enumerate(segments_per_utt):
        if segs is None:
            continue
        for unit, s, e in segs:
            for r, (a, b) in enumerate(split_segment(e - s)):
                sets[unit * E + r].extend(range(int(frame_off[u]) + s + a, int(frame_off[u]) + s + b))
    return [np.array(x, dtype=np.int64) for x in sets]


# --------------------------------------------------------------------------- stand-alone GMM EM (§8 f2)
def gmm_em(data, mean, var, alpha, c_covariance=1e-3, max_iter=10 ** 6):
    """Clustering.GMM.em(smem=False) (Clustering.py:695-719) in linear fp64 arithmetic:
    expectation (:583-600) - posteriors from the reference's log-Gaussian (Q1: -1/2 sum(var));
    maximization (:619-651) - mean, variance around the NEW mean floored at c_covariance,
    alpha = occupancy / n; q_function (:602-613) with the new parameters; loop while Q grows by
    more than 1.28.  Returns (mean, var, alpha, iterations, q_value); `gmm_em.margin` afterwards holds
    the smallest distance of a Q increment from the 1.28 threshold (tests skip knife-edge cases)."""
    X = np.asarray(data, dtype=np.float64)
    mean, var, alpha = (np.array(a, dtype=np.float64) for a in (mean, var, alpha))
    n, D = X.shape

    def log_density(mean, var):  # [n, M]
        d = X[:, None, :] - mean[None]
        return -0.5 * D * LOG_2PI - 0.5 * var.sum(-1)[None] - 0.5 * (d * d / var[None]).sum(-1)

    q_value, iters = -np.inf, 0
    gmm_em.margin = np.inf
    while iters < max_iter:
        iters += 1
        lg = log_density(mean, var) + np.log(alpha)[None]
        gam = np.exp(lg - lse(lg, axis=1, keepdims=True))
        occ = gam.sum(0)
        mean = (gam.T @ X) / occ[:, None]
        d = X[:, None, :] - mean[None]
        var = np.einsum("nm,nmd->md", gam, d * d) / occ[:, None]
        var = np.where(var < c_covariance, c_covariance, var)
        alpha = occ / n
        q = float((occ * np.log(alpha)).sum() + (gam * log_density(mean, var)).sum())
        gmm_em.margin = min(gmm_em.margin, abs(q - q_value - 1.28))
        if q - q_value > 1.28:
            q_value = q
        else:
            break
    return mean, var, alpha, iters, q_value


# ---------------------------------------------------------------------------------- front end (section 8 f4)
def mfcc_features(samples, rate, vec_num=13, sampletime=0.025, overlap=0.5, nfft=512, cal_energy=True, d1=False,
                  d2=False, filterbanks=26):
    """AudioProcessing.MFCC.mfcc (AudioProcessing.py:416-448) on an int16 / float sample array, stage by stage with
    the reference's arithmetic: pre-emphasis shifted by one sample with a zero appended (:195-198), frames of
    int(rate * sampletime) samples every int(framesize * overlap) (:215-227), ONE Hamming factor per frame computed
    from the frame index (:243-246), |rfft(frame, nfft)| (:262-263), energy = sum of the magnitudes (:338), filter
    responses with two rising flanks (:318-326), log, DCT with the (2k - 1) argument and 2 / sqrt(n) on every
    coefficient (:355-368), coefficient 0 replaced by log energy, deltas over +-2 frames with edge padding (:405-412)."""
    x = np.asarray(samples, dtype=np.float64)
    y = np.append(x[1:] - 0.98 * x[:-1], 0.0)
    framesize = int(rate * sampletime)
    step = int(framesize * overlap)
    framenum = 1 + math.ceil((len(y) - framesize) / step)
    pad = (framenum - 1) * step + framesize
    y = np.concatenate([y, np.zeros(int(pad - len(y)))])
    frames = np.stack([y[f * step:f * step + framesize] for f in range(framenum)])
    win = np.array([0.54 - 0.46 * math.cos(2 * math.pi * f / (framenum - 1)) for f in range(framenum)])
    spec = np.absolute(np.fft.rfft(frames * win[:, None], nfft))
    mel = np.linspace(0.0, 2595 * math.log(1 + (rate / 2) / 700), filterbanks + 2)
    edge = np.floor((nfft + 1) / rate * (700 * (np.exp(mel / 2595) - 1)))
    resp = np.zeros((filterbanks, nfft // 2 + 1))
    for m in range(filterbanks):
        a, b, c = int(edge[m]), int(edge[m + 1]), int(edge[m + 2])
        resp[m, a:b] = (np.arange(a, b) - a) / (edge[m + 1] - edge[m])
        resp[m, b:c] = (np.arange(b, c) - b) / (edge[m + 2] - edge[m + 1])
    with np.errstate(divide="ignore"):
        logfb = np.log(spec @ resp.T)
        energy = np.log(spec.sum(axis=1))
    k = np.arange(filterbanks)
    basis = np.cos(np.pi * (2 * k[None, :] - 1) * np.arange(vec_num)[:, None] / (2 * filterbanks))
    feat = (2 / filterbanks ** 0.5) * logfb @ basis.T
    if cal_energy:
        feat[:, 0] = energy

    def delta(f):
        p = np.pad(f, ((2, 2), (0, 0)), mode="edge")
        return sum(i * p[2 + i:2 + i + len(f)] for i in range(-2, 3)) / 10.0

    out = [feat]
    if d1:
        out.append(delta(feat))
        if d2:
            out.append(delta(out[-1]))
    return np.concatenate(out, axis=1)


def vad_filter(feat, sample=16, alpha=0.5, beta=0.93):
    """AudioProcessing.VAD.mfcc (AudioProcessing.py:462-541): (distances, filtered distances, kept frames)."""
    feat = np.asarray(feat, dtype=np.float64)
    noise = 1.0 / sample * feat[:sample].sum(axis=0)
    for i in range(sample):
        noise = alpha * noise + (1 - alpha) * feat[i]
    dist = np.sqrt(((noise[None] - feat) ** 2).sum(axis=1))
    osf = dist.copy()
    h = int(beta * (2 * sample + 1))
    for i in range(sample, len(feat) - sample):
        w = np.sort(dist[i - sample:i + sample])
        osf[i] = (1 - beta) * w[h] + beta * w[h + 1]
    thr = osf[int(sample / 2)] * (osf.max() - osf.min()) / osf.max()
    return dist, osf, feat[osf - thr > 0.0]
