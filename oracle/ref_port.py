"""CPU oracle — loop-faithful restatement of Poccala's E-step hot path (numpy fp64).

TEST INFRASTRUCTURE ONLY.  Nothing under ``poccala_b200/`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs use it, and there only as the checker / the timed CPU baseline, never as the product.

Parity status: the reference ships NO tests and NO golden vectors, so parity is
"unpinned by the reference's own tests".  The pin used instead is the *executed reference*:
``tests/golden/make_golden.py`` imports ``/root/reference`` (Appendix-B recipe), runs its own
code on seeded inputs and commits the outputs as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against those files.

This port keeps the reference's *cost structure* (one Python call per frame per mixture, dense
O(T*N^2) forward/backward, log-domain accumulators) so that timing it is a fair stand-in for
the reference CPU path on a box where ``/root/reference`` does not exist.  The state is held in
plain dicts/arrays rather than the reference's classes.

Each function cites the reference lines it restates (paths relative to /root/reference).
"""
from __future__ import annotations

import math
import random
import sys

import numpy as np

LOG_2PI = np.log(2 * math.pi)  # StatisticalModel/util.py:14
BIAS = 100.0  # StatisticalModel/Clustering.py:103
BW_THRESHOLD = 0.64  # StatisticalModel/LHMM.py:539

np.seterr(all="ignore")  # the reference silences numpy warnings globally (util.py:27)


# --------------------------------------------------------------------------- numeric helpers
def log_gaussian_diag(y, mean, var_diag, dim):
    """util.py:20-36, log branch.  Q1: the normaliser uses -0.5*sum(sigma^2), not log-det."""
    x = y - mean
    func = -dim / 2 * LOG_2PI - 0.5 * np.sum(var_diag)
    quad = -0.5 * np.dot(x * (1.0 / var_diag), x)
    return func + quad


def lse(v):
    """util.py:54-77 (scalar form): max-shifted log-sum-exp; +-inf maximum is returned as is."""
    v = np.asarray(v)
    mx = np.max(v)
    if abs(mx) == float("inf"):
        return mx
    return mx + np.log(np.sum(np.exp(v - mx)))


def lse_rows(a):
    """util.py:70-75: log_sum_exp(..., vector=True) - one lse per leading index."""
    return np.array([lse(a[i]) for i in range(len(a))])


def lse_stack(arrays, rows):
    """util.py:80-92 matrix_log_sum_exp: element-wise lse over a list of [rows, cols] arrays."""
    out = []
    for i in range(rows):
        block = np.array([a[i, :] for a in arrays]).T
        out.append(lse_rows(block))
    return np.array(out)


# --------------------------------------------------------------------------- model containers
def new_unit(state_num, mix, dim):
    """AcousticModel.py:164-226 init_unit: 5-state left-to-right HMM, entry row (0,1,0..),
    emitting rows (0.5 self, 0.5 next), exit row zero; log-domain accumulators start at -inf
    (LHMM.py:84-85, Clustering.py:98-101)."""
    A = np.zeros((state_num, state_num))
    A[0, 1] = 1.0
    for j in range(1, state_num - 1):
        A[j, j] = 0.5
        A[j, j + 1] = 0.5
    E = state_num - 2
    gmms = []
    for _ in range(E):
        gmms.append(
            dict(
                alpha=np.ones(mix) / mix,
                mean=np.zeros((mix, dim)),
                var=np.ones((mix, dim)),
                acc=np.full(mix, -np.inf),
                alpha_acc=-np.inf,
                mean_acc=np.full((mix, dim), -np.inf),
                cov_acc=np.full((mix, dim), -np.inf),
                record=[],
            )
        )
    return dict(
        transmat=A,
        gmms=gmms,
        state_num=state_num,
        ksai_acc=np.full((E, state_num), -np.inf),
        gamma_acc=np.full(E, -np.inf),
        B=None,
    )


def clone_unit(u):
    """Fresh per-label-position instance carrying the same parameters (the reference builds a
    new LHMM + GMMs per label position and loads parameters from disk: AcousticModel.py:897-902)."""
    S = u["state_num"]
    mix, dim = u["gmms"][0]["mean"].shape
    n = new_unit(S, mix, dim)
    n["transmat"] = u["transmat"].copy()
    for g, s in zip(n["gmms"], u["gmms"]):
        g["alpha"] = np.array(s["alpha"], dtype=np.float64).copy()
        g["mean"] = s["mean"].copy()
        g["var"] = s["var"].copy()
    return n


# --------------------------------------------------------------------------- scoring (a1-a3)
def gmm_point(g, x, record=True):
    """Clustering.py:740-767 point(log=True): per-mixture log alpha + log N, recorded, then lse."""
    mix, dim = g["mean"].shape
    if len(x) != dim:
        raise ValueError("DataDimensionError: expected %d got %d" % (dim, len(x)))  # :749-751
    p_list = []
    for i in range(mix):
        p_list.append(np.log(g["alpha"][i]) + log_gaussian_diag(x, g["mean"][i], g["var"][i], dim))
    if record:
        g["record"].append(p_list)
    return lse(p_list)


def cal_observation_pro(unit, X):
    """LHMM.py:163-187: B[S, T]; row 0 = VirtualState(1.) -> log 1 = 0, row S-1 = VirtualState(0.)
    -> -inf (AcousticModel.py:218-222, 1039-1043)."""
    S = unit["state_num"]
    T = len(X)
    rows = [[np.log(1.0) for _ in range(T)]]
    for g in unit["gmms"]:
        rows.append([gmm_point(g, X[t], record=True) for t in range(T)])
    rows.append([np.log(0.0) for _ in range(T)])
    unit["B"] = np.array(rows)
    assert unit["B"].shape == (S, T)
    return unit["B"]


# --------------------------------------------------------------------------- sentence HMM (a4)
def embedded(label, hmm_list, state_num):
    """AcousticModel.py:957-1014: states dict, pasted A[N,N], stacked B[N,T], uniform pi (Q4)."""
    E = state_num - 2
    L = len(label)
    N = E * L + 2
    states = {0: label[0]}
    k = 1
    for u in label:
        for _ in range(E):
            states[k] = u
            k += 1
    states[k] = label[-1]
    A = np.zeros((N, N))
    A[: state_num - 1, :state_num] = hmm_list[0]["transmat"][:-1]  # :981
    for i in range(L):
        a = i * E + 1
        b = (i + 1) * E + 1
        A[a:b, a - 1 : a - 1 + state_num] = hmm_list[i]["transmat"][1:-1]  # :987
    B = hmm_list[0]["B"][0:-1]
    for i in range(1, L):
        B = np.append(B, hmm_list[i]["B"][1:-1, :], axis=0)
    B = np.append(B, hmm_list[L - 1]["B"][-1:, :], axis=0)
    pi = np.ones(N) / N
    return states, A, B, pi


# --------------------------------------------------------------------------- forward/backward
def forward_dense(A, B, pi):
    """LHMM.py:335-351: dense log-domain forward, one lse per (t, j)."""
    N, T = B.shape
    logA = np.log(A)
    f = np.zeros((N, T))
    f[:, 0] = np.log(pi) + B[:, 0]
    for t in range(1, T):
        p = []
        for j in range(N):
            p.append(lse(f[:, t - 1] + logA[:, j]))
        f[:, t] = np.array(p) + B[:, t]
    return f


def backward_dense(A, B):
    """LHMM.py:353-366: beta[:, T-1] = 0 (buffer is zero-initialised :383), dense recursion."""
    N, T = B.shape
    logA = np.log(A)
    b = np.zeros((N, T))
    for t in range(T - 2, -1, -1):
        rows = []
        for j in range(N):
            rows.append(logA[j, :] + B[:, t + 1] + b[:, t + 1])
        b[:, t] = lse_rows(rows)
    return b


def maximization(A, B, f, b):
    """LHMM.py:426-471 (single-utterance branch): unnormalised ksai/gamma (Q6) and new pi."""
    N, T = B.shape
    logA = np.log(A)
    per_t = []
    for t in range(T - 1):
        m_t = []
        for m in range(N):
            m_t.append(f[m, t] + logA[m] + B[:, t + 1] + b[:, t + 1])  # :394-405
        per_t.append(np.array(m_t))
    if per_t:
        ksai = lse_stack(per_t, N)
    else:  # T == 1: nothing to sum
        ksai = np.full((N, N), -np.inf)
    gamma = lse_rows(f[:, :-1] + b[:, :-1]) if T > 1 else np.full(N, -np.inf)
    pi_arr = f[:, 0] + b[:, 0]
    pi_arr = pi_arr - lse(pi_arr)
    return ksai, gamma, np.exp(pi_arr)


def baum_welch_utterance(A, B, pi, max_iter=1000):
    """LHMM.py:526-544: loop {forward, backward, maximization, expectation} while the
    log-likelihood grows by more than 0.64; only pi changes between iterations (Q5).
    Returns the alpha/beta/ksai/gamma of the LAST executed iteration plus bookkeeping."""
    q = -float("inf")
    iters = 0
    while True:
        iters += 1
        f = forward_dense(A, B, pi)
        b = backward_dense(A, B)
        ksai, gamma, new_pi = maximization(A, B, f, b)
        q_new = lse(f[:, -1])  # LHMM.py:412-422
        pi_used = pi
        pi = new_pi  # fix_code pi-bit is 0 on this path
        if q_new - q > BW_THRESHOLD and iters < max_iter:
            q = q_new
        else:
            return dict(alpha=f, beta=b, ksai=ksai, gamma=gamma, pi_used=pi_used, pi_next=pi,
                        logp=q_new, iters=iters)


# --------------------------------------------------------------------------- accumulators
def hmm_add_acc(unit, ksai_value, gamma_value):
    """LHMM.py:149-161."""
    E = unit["state_num"] - 2
    unit["ksai_acc"] = lse_stack([unit["ksai_acc"], ksai_value], E)
    unit["gamma_acc"] = lse_stack([unit["gamma_acc"].reshape(1, -1), gamma_value.reshape(1, -1)], 1).reshape(-1)


def gmm_update_acc(g, l_value, b_value, o_value):
    """Clustering.py:653-680: log-domain occupancy / (x+100) / (x-mu_old)^2 accumulators (Q7, Q8)."""
    rec = np.array(g["record"]).T  # [M, T]
    rec = rec + (l_value - b_value)
    o_t = o_value.T
    log_o = np.log(o_t + BIAS)
    mix = rec.shape[0]
    g["acc"] = lse_rows(np.append(rec, g["acc"].reshape(-1, 1), axis=1))
    g["alpha_acc"] = lse(np.append(l_value, g["alpha_acc"]))
    for m in range(mix):
        g["mean_acc"][m] = lse_rows(np.append(log_o + rec[m], g["mean_acc"][m].reshape(-1, 1), axis=1))
    for m in range(mix):
        g["cov_acc"][m] = lse_rows(
            np.append(rec[m] + np.log((o_t - g["mean"][m].reshape(-1, 1)) ** 2), g["cov_acc"][m].reshape(-1, 1), axis=1)
        )
    g["record"] = []


def update_acc(bw, B, X, hmm_list, state_num, fix_code=0):
    """LHMM.py:473-507: slice ksai/gamma per label position; per-frame normalised log-gamma rows
    feed the GMM accumulators."""
    E = state_num - 2
    fix_trans = bool(fix_code & 4)
    fix_pdf = bool(fix_code & 2)
    ksai_view = bw["ksai"][1:-1, :]
    gamma_view = bw["gamma"][1:-1]
    l_value = bw["alpha"] + bw["beta"]
    sum_value = lse_rows(l_value.T)
    l_value = l_value[1:-1]
    b_value = B[1:-1, :]
    ix = iy = 0
    lgam = np.empty_like(l_value)
    for hmm in hmm_list:
        if not fix_trans:
            hmm_add_acc(hmm, ksai_view[iy : iy + E, ix : ix + state_num], gamma_view[iy : iy + E])
        lrows = l_value[iy : iy + E, :] - sum_value
        lgam[iy : iy + E, :] = lrows
        if not fix_pdf:
            for i in range(E):
                gmm_update_acc(hmm["gmms"][i], lrows[i, :], b_value[iy + i, :], X)
        iy += E
        ix += E
    return lgam


def estep_utterance(units, label, X, state_num=5, fix_code=0, keep=False):
    """AcousticModel.py:884-916 multi_embedded_training_1 without the file IO: score every label
    position with a fresh copy of its unit, build the sentence HMM, run baulm_welch, return the
    per-position accumulators (what __save_acc would have written)."""
    X = np.asarray(X, dtype=np.float64)
    hmm_list = []
    for u in label:
        h = clone_unit(units[u])
        cal_observation_pro(h, X)
        hmm_list.append(h)
    states, A, B, pi = embedded(label, hmm_list, state_num)
    bw = baum_welch_utterance(A, B, pi)
    lgam = update_acc(bw, B, X, hmm_list, state_num, fix_code)
    out = dict(hmm_list=hmm_list, logp=bw["logp"], iters=bw["iters"])
    if keep:
        out.update(A=A, B=B, pi=pi, states=states, bw=bw, lgam=lgam)
    return out


def merge_accs(units, per_position):
    """LHMM.py:256-290 + Clustering.py:314-367 init_acc: lse-merge of every saved accumulator of a
    unit (``per_position`` = list of (unit_name, hmm_dict) in save order; loss-free, i.e. without
    the same-second filename collision Q13)."""
    merged = {}
    for name, u in units.items():
        mine = [h for (n, h) in per_position if n == name]
        if not mine:
            continue
        m = clone_unit(u)
        E = u["state_num"] - 2
        m["ksai_acc"] = lse_stack([h["ksai_acc"] for h in mine], E)
        m["gamma_acc"] = lse_rows(np.array([h["gamma_acc"] for h in mine]).T)
        for gi, g in enumerate(m["gmms"]):
            g["acc"] = lse_rows(np.array([g["acc"]] + [h["gmms"][gi]["acc"] for h in mine]).T)
            g["alpha_acc"] = lse(np.array([g["alpha_acc"]] + [h["gmms"][gi]["alpha_acc"] for h in mine]))
            stackm = np.array([g["mean_acc"]] + [h["gmms"][gi]["mean_acc"] for h in mine])
            stackc = np.array([g["cov_acc"]] + [h["gmms"][gi]["cov_acc"] for h in mine])
            mix, dim = g["mean"].shape
            for k in range(mix):
                g["mean_acc"][k] = lse_rows(stackm[:, k, :].T)
                g["cov_acc"][k] = lse_rows(stackc[:, k, :].T)
        merged[name] = m
    return merged


def update_param(unit, c_covariance=1e-3, fix_code=0):
    """LHMM.py:509-524 + Clustering.py:682-693."""
    S = unit["state_num"]
    if not (fix_code & 4):
        unit["transmat"][1:-1, :] = np.exp(unit["ksai_acc"] - unit["gamma_acc"].reshape((S - 2, 1)))
    if not (fix_code & 2):
        for g in unit["gmms"]:
            g["alpha"] = np.exp(g["acc"] - g["alpha_acc"])
            g["mean"] = np.exp(g["mean_acc"] - g["acc"].reshape(-1, 1)) - BIAS
            for m in range(len(g["alpha"])):
                c = np.exp(g["cov_acc"][m] - g["acc"][m])
                c[np.where(c < c_covariance)] = c_covariance
                g["var"][m] = c
    return unit


def em_iteration(units, utterances, state_num=5, c_covariance=1e-3, fix_code=0):
    """AcousticModel.py:842-882 embedded_training: E-step over utterances, then M-step per unit.
    ``utterances`` = list of (label, X).  Returns (new_units, total_logp)."""
    saved = []
    total = 0.0
    for label, X in utterances:
        r = estep_utterance(units, label, X, state_num, fix_code)
        for name, h in zip(label, r["hmm_list"]):
            saved.append((name, h))
        total += r["logp"]
    merged = merge_accs(units, saved)
    new_units = {}
    for name, u in units.items():
        if name in merged:
            new_units[name] = update_param(merged[name], c_covariance, fix_code)
        else:
            new_units[name] = clone_unit(u)
    return new_units, total


# --------------------------------------------------------------------------- Viterbi (a16)
def viterbi_dense(A, prob, pi):
    """LHMM.py:546-609 (end_state_back=False): fp64 max-plus, first-argmax ties, end state =
    first argmax over all states.  Returns (score, float64 state path)."""
    N, T = prob.shape
    mark = np.zeros((T,))
    before = [[0 for _ in range(T)] for _ in range(N)]
    p = np.log(pi) + prob[:, 0]
    logA = np.log(A)
    max_index = 0
    for t in range(1, T):
        p_ = np.zeros_like(p)
        for j in range(N):
            tmp = p + logA[:, j]
            mx = tmp.max()
            p_[j] = mx
            before[j][t] = np.where(tmp == mx)[0][0]
        p = p_ + prob[:, t]
    max_index = np.where(p == p.max())[0][0]
    score = p[max_index]
    idx = max_index
    for t in range(T - 1, -1, -1):
        mark[t] = idx
        idx = before[idx][t]
    return score, mark


# --------------------------------------------------------------------------- k-means (a17)
def cal_distance(d1, d2, arg=2):
    """Clustering.py:796-801 - Q2: the return sits inside the loop, so only dimension 0 counts."""
    dist = 0.0
    for i in range(len(d1)):
        dist += abs(d1[i] - d2[i]) ** arg
        return dist ** (1 / arg)


def cal_variance_kmeans(center, points):
    """Clustering.py:807-832 (algorithm='kmeans'): per-dimension std with a 1e-4 variance floor."""
    out = []
    for d in range(len(center)):
        v = 0.0
        for p in points:
            v += (center[d] - p[d]) ** 2
        v /= len(points)
        if v < 1e-4:
            v = 1e-4
        out.append(v ** 0.5)
    return out


def kmeans_pp(data, k, cov_matrix=True, trace=None):
    """Clustering.py:838-1044, algorithm=1: first-seed-only k-means++ weights (Q9), greedy
    one-point-per-cluster passes that abort on the first empty-handed cluster (Q10), seeds never
    leave their cluster (Q11).  Uses the *global* ``random`` module exactly like the reference:
    call ``random.seed(s)`` first.  Returns (mean[K,D], cov[K,D,D] or std[K,D], alpha, clusters)."""
    n = len(data)
    mark = [[-1, -1] for _ in range(n)]
    cl = []  # [centre, {key: point}]
    c0 = random.randint(0, n - 1)
    cl.append([np.array(data[c0], dtype=np.float64, copy=True), {-1: data[c0]}])
    mark[c0][0] = 0
    total = 0.0
    dl = []
    for d in data:
        mind = sys.maxsize
        for c in cl:
            dist = cal_distance(c[0], d)
            if dist < mind:
                mind = dist
        dl.append(mind)
        total += mind
    if total == 0.0:
        idx = random.sample(range(0, len(dl)), k - 1)
        for kk in range(1, k):
            for i in idx:
                cl.append([np.array(data[i], dtype=np.float64, copy=True), {-1: data[i]}])
                mark[i][0] = kk
        assert len(cl) == k
    else:
        for kk in range(1, k):
            r = random.randint(0, int(total))
            for i in range(len(dl)):
                r -= dl[i]
                if r < 0:
                    cl.append([np.array(data[i], dtype=np.float64, copy=True), {-1: data[i]}])
                    mark[i][0] = kk
                    break
        assert len(cl) == k
    moved = True
    passes = 0
    while moved:
        moved = False
        passes += 1
        for kk in range(k):
            best = sys.maxsize
            bi = -1
            move_class = False
            for i in range(n):
                if mark[i][0] == kk:
                    continue
                dist = cal_distance(cl[kk][0], data[i])
                if mark[i][0] != -1:
                    dist_own = cal_distance(cl[mark[i][0]][0], data[i])
                    if dist_own <= dist:
                        continue
                if dist < best:
                    best = dist
                    bi = i
                    move_class = mark[i][0] != -1 and mark[i][1] != -1
            if bi == -1:
                break
            moved = True
            if move_class:
                del cl[mark[bi][0]][1][mark[bi][1]]
            key = int(random.random() * 1e15)
            while cl[kk][1].get(key):
                key = int(random.random() * 1e15)
            mark[bi][1] = key
            mark[bi][0] = kk
            cl[kk][1][key] = data[bi]
            if trace is not None:
                trace.append((passes, kk, bi))
        for kk in range(k):
            c = np.copy(cl[kk][0])
            pts = cl[kk][1]
            for d in range(len(c)):
                c[d] = 0.0
                for p in pts.values():
                    c[d] += p[d]
                c[d] /= len(pts)
            cl[kk][0] = c
    means = np.array([cl[i][0] for i in range(k)])
    stds = np.array([cal_variance_kmeans(cl[i][0], cl[i][1].values()) for i in range(k)])
    alpha = [len(cl[i][1]) / n for i in range(k)]
    clusters = [list(cl[i][1].values()) for i in range(k)]
    if cov_matrix:
        stds = np.array([np.diag(s ** 2) for s in stds])
    return means, stds, alpha, clusters
