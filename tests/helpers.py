"""Shared helpers for the parity tests."""
import os

import numpy as np

UNITS3 = ["a", "b", "c"]


def load_golden(name):
    here = os.path.dirname(os.path.abspath(__file__))
    return np.load(os.path.join(here, "golden", name), allow_pickle=False)


def maxdiff(a, b):
    """Max |a-b| over finite entries; the non-finite patterns must agree exactly."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    fa, fb = np.isfinite(a), np.isfinite(b)
    assert (fa == fb).all(), "finite pattern differs"
    assert (a[~fa] == b[~fb]).all() or (np.isnan(a[~fa]) == np.isnan(b[~fb])).all()
    if not fa.any():
        return 0.0
    return float(np.abs(a[fa] - b[fb]).max())


def port_units(mean, var, alpha, names=UNITS3, transmat=None):
    from oracle import ref_port as rp

    units = {}
    for i, u in enumerate(names):
        d = rp.new_unit(5, mean.shape[2], mean.shape[3])
        for r, g in enumerate(d["gmms"]):
            g["mean"] = np.array(mean[i, r], dtype=np.float64)
            g["var"] = np.array(var[i, r], dtype=np.float64)
            g["alpha"] = np.array(alpha[i, r], dtype=np.float64)
        if transmat is not None:
            d["transmat"] = np.array(transmat[i], dtype=np.float64)
        units[u] = d
    return units


def alignment_case(g, tag):
    """One case of tests/golden/alignment.npz: (labels, lens, sequences or None, {key: frame set})."""
    n = int(g[f"{tag}_n_utt"])
    labels = [g[f"{tag}_label{u}"].astype(np.int32) for u in range(n)]
    lens = g[f"{tag}_lens"].astype(np.int32)
    seqs = [g[f"{tag}_seq{u}"] for u in range(n)] if f"{tag}_seq0" in g else None
    sets = {int(k[len(tag) + 5:]): g[k] for k in g if k.startswith(tag + "_set_")}
    return labels, lens, seqs, sets
