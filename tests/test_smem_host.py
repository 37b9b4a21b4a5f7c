"""§8 f2, host half of the split / merge search (Clustering.py:371-430): the merge ranking (cosine of `__gamma` rows) and
the split ranking (rank-weighted posteriors against the `standard=True` density) computed from the SAME tables the executed
reference held when its __SMEM started (tests/golden/gmm_smem.npz: gamma_in = log posteriors of the last expectation(),
the current parameters).  Pure numpy on the host - no device needed; the device half (the tables themselves, the candidate's
re-estimation and the Q terms) is tests/test_gpu_surface.py."""
import numpy as np
import pytest

from tests.helpers import load_golden

torch = pytest.importorskip("torch")


def test_merge_and_split_rankings_match_executed_reference():
    from poccala_b200.Clustering import Clustering

    g = load_golden("gmm_smem.npz")
    for c in range(int(g["n"])):
        gamma, data = g[f"s{c}_gamma_in"], g[f"s{c}_data"]
        M = gamma.shape[0]
        gm = Clustering.GMM(None, dimension=39, mix_level=M)
        merge = gm._j_merge(gamma)
        want = g[f"s{c}_merge"]
        assert [(a, b) for a, b, _ in merge] == [(int(a), int(b)) for a, b, _ in want]
        assert np.allclose([v for _, _, v in merge], want[:, 2], rtol=1e-12, atol=0)
        split, owner = gm._j_split(gamma, data, g[f"s{c}_mean_in"], g[f"s{c}_var_in"])
        want = g[f"s{c}_split"]
        assert [k for k, _ in split] == [int(k) for k in want[:, 0]]
        got = np.array([v for _, v in split])
        assert np.array_equal(np.isnan(got), np.isnan(want[:, 1]))
        ok = ~np.isnan(got)
        assert np.allclose(got[ok], want[ok, 1], rtol=1e-10, atol=0)
        assert np.array_equal(owner, np.argmax(gamma, axis=0))
