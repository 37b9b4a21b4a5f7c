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
