"""Deterministic synthetic 39-dim MFCC-like workloads.

Hierarchical ground truth so that mixture posteriors are non-degenerate: per (unit, state) a
centre ~ N(0, 1.5^2), component means = centre + N(0, 0.7^2), variances ~ U(0.5, 1.5),
weights ~ Dirichlet(2).  Utterances follow a label sequence; every emitting state gets a run of
frames drawn from its GMM.  |x| stays << 100 (the reference's +100 mean bias, Q7).

numpy generators are the canonical ones for parity tests; ``torch_*`` variants build
bench-scale batches directly in HBM.
"""
from __future__ import annotations

import os

import numpy as np

DIM = 39
STATE_NUM = 5
EMIT = STATE_NUM - 2

# AcousticModel/Unit/IF of the reference (data, not code): 22 initials + 35 finals = 57 units.
IF_INITIALS = "b,p,m,f,d,t,n,l,g,k,h,j,q,x,zh,ch,sh,z,c,s,r,#".split(",")
IF_FINALS = ("a,e,i,u,an,ao,en,ing,iang,iao,o,ai,in,ou,uo,vn,ong,v,ei,ia,ie,iu,ua,ui,un,ve,eng,ian,uai,uan,"
             "van,uang,ang,er,iong").split(",")
IF_UNITS = IF_INITIALS + IF_FINALS


def load_unit_file(path):
    """Same file format the reference parses in AcousticModel.load_unit (AcousticModel.py:134-162):
    first line is a title, the remaining lines are comma-separated unit names."""
    units = []
    with open(path) as f:
        f.readline()
        for line in f:
            line = line.strip("\n")
            if line:
                units.extend(line.split(","))
    return units


def make_truth(n_units, mix, seed, dim=DIM):
    """Ground-truth parameters: mean[n_units,3,M,D], var[...], alpha[n_units,3,M]."""
    rng = np.random.default_rng(seed)
    centre = rng.normal(0.0, 1.5, size=(n_units, EMIT, 1, dim))
    mean = centre + rng.normal(0.0, 0.7, size=(n_units, EMIT, mix, dim))
    var = rng.uniform(0.5, 1.5, size=(n_units, EMIT, mix, dim))
    alpha = rng.dirichlet(np.full(mix, 2.0), size=(n_units, EMIT))
    return mean, var, alpha


def perturb(mean, var, alpha, seed):
    """Initial model = truth perturbed by N(0,0.3^2) on means and xU(0.8,1.25) on variances."""
    rng = np.random.default_rng(seed)
    return (mean + rng.normal(0.0, 0.3, size=mean.shape), var * rng.uniform(0.8, 1.25, size=var.shape),
            alpha.copy())


def default_transmat(n_units):
    """AcousticModel.init_unit's initial transition matrix (AcousticModel.py:176-181)."""
    A = np.zeros((STATE_NUM, STATE_NUM))
    A[0, 1] = 1.0
    for j in range(1, STATE_NUM - 1):
        A[j, j] = 0.5
        A[j, j + 1] = 0.5
    return np.repeat(A[None], n_units, axis=0)


def random_labels(n_utt, L, n_units, seed, n_initials=None):
    """Label matrix [n_utt, L] of unit indices.  With ``n_initials`` set, alternate
    initial/final like Mandarin syllables (SURVEY §8d fallback generator)."""
    rng = np.random.default_rng(seed)
    if n_initials is None or n_initials <= 0 or n_initials >= n_units:
        return rng.integers(0, n_units, size=(n_utt, L), dtype=np.int32)
    lab = np.empty((n_utt, L), dtype=np.int32)
    lab[:, 0::2] = rng.integers(0, n_initials, size=lab[:, 0::2].shape)
    lab[:, 1::2] = rng.integers(n_initials, n_units, size=lab[:, 1::2].shape)
    return lab


def state_durations(rng, n_states, T):
    """Split T frames over n_states runs, each >= 1 when T allows."""
    if T >= n_states:
        cuts = np.sort(rng.choice(np.arange(1, T), size=n_states - 1, replace=False)) if n_states > 1 else np.array([], int)
        edges = np.concatenate([[0], cuts, [T]])
        return np.diff(edges)
    d = np.zeros(n_states, dtype=np.int64)
    d[:T] = 1
    return d


def make_utterance(label, T, truth, seed):
    """One utterance X[T, D] (fp64) following ``label`` (sequence of unit indices)."""
    mean, var, alpha = truth
    rng = np.random.default_rng(seed)
    n_states = EMIT * len(label)
    dur = state_durations(rng, n_states, T)
    mix = mean.shape[2]
    rows = []
    for s, d in enumerate(dur):
        if d == 0:
            continue
        u, r = label[s // EMIT], s % EMIT
        comp = rng.choice(mix, size=d, p=alpha[u, r])
        rows.append(mean[u, r, comp] + rng.normal(size=(d, mean.shape[-1])) * np.sqrt(var[u, r, comp]))
    return np.concatenate(rows, axis=0)


def make_corpus(n_utt, T, L, n_units, mix, cfg_seed, n_initials=None, ragged=False):
    """Corpus for a config: (truth, init params, labels[list of arrays], utterances[list of X])."""
    truth = make_truth(n_units, mix, 1000 * cfg_seed + 7)
    init = perturb(*truth, seed=1000 * cfg_seed + 11)
    rng = np.random.default_rng(1000 * cfg_seed + 13)
    labels, utts = [], []
    base = random_labels(n_utt, L, n_units, 1000 * cfg_seed + 17, n_initials)
    for i in range(n_utt):
        Li = int(rng.integers(1, L + 1)) if ragged else L
        Ti = int(rng.integers(max(4, T // 3), T + 1)) if ragged else T
        lab = base[i, :Li].copy()
        labels.append(lab)
        utts.append(make_utterance(lab, Ti, truth, 1000 * cfg_seed + 100 + i))
    return truth, init, labels, utts


# --------------------------------------------------------------------------- torch (bench scale)
def torch_corpus(n_utt, T, L, n_units, mix, seed, device, n_initials=None, dim=DIM, data_seed=None):
    """Fixed-shape corpus generated on ``device``: returns (truth, init, labels[n_utt,L] int32 numpy,
    X[n_utt*T, D] float32 device tensor).  Durations are near-uniform with jitter; frames are drawn
    from the owning state's GMM.  Same hierarchical model as the numpy generator."""
    import torch

    truth = make_truth(n_units, mix, 1000 * seed + 7, dim)
    init = perturb(*truth, seed=1000 * seed + 11)
    data_seed = seed if data_seed is None else data_seed  # same model, different utterances per rank
    labels = random_labels(n_utt, L, n_units, 1000 * data_seed + 17, n_initials)
    g = torch.Generator(device=device)
    g.manual_seed(1000 * data_seed + 19)
    n_states = EMIT * L
    mean_t = torch.as_tensor(truth[0], dtype=torch.float32, device=device).reshape(n_units * EMIT, mix, dim)
    std_t = torch.as_tensor(np.sqrt(truth[1]), dtype=torch.float32, device=device).reshape(n_units * EMIT, mix, dim)
    cum_alpha = torch.as_tensor(np.cumsum(truth[2], axis=-1), dtype=torch.float32, device=device).reshape(
        n_units * EMIT, mix)
    lab_t = torch.as_tensor(labels.astype(np.int64), device=device)
    X = torch.empty((n_utt * T, dim), dtype=torch.float32, device=device)
    chunk = max(1, min(n_utt, (1 << 22) // max(T, 1)))
    for lo in range(0, n_utt, chunk):
        hi = min(n_utt, lo + chunk)
        n = hi - lo
        # random monotone state index per frame: sorted uniform cut points
        cuts = torch.rand((n, n_states - 1), generator=g, device=device).sort(dim=1).values if n_states > 1 else None
        tt = (torch.arange(T, device=device, dtype=torch.float32) + 0.5) / T
        if cuts is not None:
            sidx = torch.searchsorted(cuts, tt.expand(n, T).contiguous())
        else:
            sidx = torch.zeros((n, T), dtype=torch.int64, device=device)
        unit = torch.gather(lab_t[lo:hi], 1, sidx // EMIT)
        gs = unit * EMIT + (sidx % EMIT)  # global state id [n, T]
        u01 = torch.rand((n, T), generator=g, device=device)
        comp = (u01.unsqueeze(-1) > cum_alpha[gs]).sum(-1).clamp_(max=mix - 1)
        mu = mean_t[gs, comp]
        sd = std_t[gs, comp]
        noise = torch.randn((n, T, dim), generator=g, device=device)
        X[lo * T : hi * T] = (mu + noise * sd).reshape(n * T, dim)
    return truth, init, labels, X


def reference_unit_file():
    """Path of the IF unit file shipped with this package (same content/format as the reference's
    AcousticModel/Unit/IF)."""
    return os.path.join(os.path.dirname(__file__), "Unit", "IF")
