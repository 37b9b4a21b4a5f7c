"""§8 f1: the on-disk parameter layout (LHMM.py:192-254, Clustering.py:234-312).  Files written by the
mirrors must load in the reference and vice versa.  The interchange half needs /root/reference and
runs in the build container only; the layout half runs anywhere."""
import os
import sys

import numpy as np
import pytest

REF = "/root/reference"


def _models():
    from poccala_b200.Clustering import Clustering
    from poccala_b200.LHMM import LHMM

    rng = np.random.default_rng(3)
    gmm = Clustering.GMM(None, dimension=39, mix_level=4, alpha=rng.dirichlet(np.ones(4)),
                         mean=rng.normal(size=(4, 39)), variance=rng.uniform(0.5, 2, size=(4, 39)), gmm_id=2)
    tm = np.zeros((5, 5))
    tm[0, 1] = 1
    for j in range(1, 4):
        tm[j, j], tm[j, j + 1] = 0.6, 0.4
    hmm = LHMM({i: "a" for i in range(5)}, 5, None, transmat=tm, probmat=[np.zeros((5, 1))], pi=np.ones(5) / 5,
               fix_code=2)
    return gmm, hmm


def test_layout(tmp_path):
    gmm, hmm = _models()
    gmm.save_parameter(str(tmp_path))
    hmm.save_parameter(str(tmp_path))
    assert sorted(os.listdir(tmp_path / "GMM_2")) == ["GMM_config.ini", "GMM_covariance.npy", "GMM_means.npy",
                                                     "GMM_weight.npy"]
    assert sorted(os.listdir(tmp_path / "HMM")) == ["HMM_config.ini", "pi.npy", "transmat.npy"]
    assert np.load(tmp_path / "GMM_2" / "GMM_covariance.npy").shape == (4, 39, 39)
    assert np.load(tmp_path / "GMM_2" / "GMM_means.npy").shape == (4, 39)
    from poccala_b200.Clustering import Clustering, ParameterFileExistsError

    g2 = Clustering.GMM(None, dimension=39, mix_level=4, gmm_id=2)
    g2.init_parameter(str(tmp_path))
    assert np.array_equal(g2.mean, gmm.mean) and np.array_equal(g2.covariance, gmm.covariance)
    assert np.array_equal(g2.alpha, gmm.alpha)
    with pytest.raises(ParameterFileExistsError):
        Clustering.GMM(None, dimension=39, mix_level=4, gmm_id=7).init_parameter(str(tmp_path))


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree (build container only)")
def test_interchange_with_reference(tmp_path):
    from oracle import ref_harness as rh

    H = rh.Harness(["a"], 4)
    gmm, hmm = _models()
    # ours -> reference
    gmm.save_parameter(str(tmp_path))
    hmm.save_parameter(str(tmp_path))
    rg = H.Clustering.GMM(H.log, dimension=39, mix_level=4, gmm_id=2)
    rg.init_parameter(str(tmp_path))
    assert np.array_equal(rg.mean, gmm.mean) and np.array_equal(rg.covariance, gmm.covariance)
    assert np.array_equal(rg.alpha, gmm.alpha)
    rh_unit = H.am.init_unit("a", new_log=True)
    rh_unit.init_parameter(str(tmp_path))
    assert np.array_equal(rh_unit.transmat, hmm.transmat) and np.array_equal(rh_unit.pi, hmm.pi)
    # reference -> ours
    out = tmp_path / "ref"
    os.mkdir(out)
    rg.save_parameter(str(out))
    rh_unit.save_parameter(str(out))
    from poccala_b200.Clustering import Clustering

    g2 = Clustering.GMM(None, dimension=39, mix_level=4, gmm_id=2)
    g2.init_parameter(str(out))
    assert np.array_equal(g2.mean, gmm.mean) and np.array_equal(g2.covariance, gmm.covariance)
    gmm2, hmm2 = _models()
    hmm2.change_A(np.zeros((5, 5)))
    hmm2.init_parameter(str(out))
    assert np.array_equal(hmm2.transmat, hmm.transmat)
    for name in ("GMM_2", "HMM"):
        assert sorted(os.listdir(out / name)) == sorted(os.listdir(tmp_path / name))


# ---- accumulator files (LHMM.py:211-231,256-290; Clustering.py:257-283,314-367) -----------------------------
def _fill_acc(gmm, hmm, seed):
    rng = np.random.default_rng(seed)
    M, D = gmm.mixture, gmm.dimension
    occ = rng.uniform(0.5, 20.0, size=M)
    sx = occ[:, None] * rng.normal(size=(M, D))
    gmm._occ, gmm._socc, gmm._sx = occ, float(occ.sum() * 1.01), sx
    gmm._scc = occ[:, None] * rng.uniform(0.2, 2.0, size=(M, D))
    ks = np.full((3, 5), -np.inf)
    for r in range(3):
        ks[r, r + 1], ks[r, r + 2] = rng.uniform(-2e4, -1e4, size=2)
    hmm.add_acc(ks, np.logaddexp(ks[[0, 1, 2], [1, 2, 3]], ks[[0, 1, 2], [2, 3, 4]]))


def test_accumulator_layout_and_merge(tmp_path):
    """Two saves, one load: the linear statistics add up, the transition accumulators log-add."""
    from poccala_b200.Clustering import Clustering

    gmm, hmm = _models()
    _fill_acc(gmm, hmm, 1)
    gmm.save_acc(str(tmp_path)); hmm.save_acc(str(tmp_path))
    occ1, socc1, sx1, scc1 = gmm._occ.copy(), gmm._socc, gmm._sx.copy(), gmm._scc.copy()
    ks1, ga1 = hmm.ksai_acc.copy(), hmm.gamma_acc.copy()
    gmm2, hmm2 = _models()
    _fill_acc(gmm2, hmm2, 2)
    gmm2.save_acc(str(tmp_path)); hmm2.save_acc(str(tmp_path))  # same second: the names must not collide (Q13)
    assert sorted(os.listdir(tmp_path / "GMM_2")) == ["acc", "alpha-acc", "covariance-acc", "mean-acc"]
    assert all(len(os.listdir(tmp_path / "GMM_2" / d)) == 2 for d in ("acc", "alpha-acc", "covariance-acc", "mean-acc"))
    assert sorted(os.listdir(tmp_path / "HMM")) == ["gamma-acc", "ksai-acc"]
    assert np.load(tmp_path / "GMM_2" / "mean-acc" / sorted(os.listdir(tmp_path / "GMM_2" / "mean-acc"))[0]).shape == (4, 39)
    g3, h3 = _models()
    g3.init_acc(str(tmp_path)); h3.init_acc(str(tmp_path))
    assert np.allclose(g3._occ, occ1 + gmm2._occ, rtol=1e-12)
    assert np.isclose(g3._socc, socc1 + gmm2._socc, rtol=1e-12)
    assert np.allclose(g3._sx, sx1 + gmm2._sx, rtol=1e-9, atol=1e-9)
    assert np.allclose(g3._scc, scc1 + gmm2._scc, rtol=1e-12)
    with np.errstate(invalid="ignore"):
        assert np.allclose(h3.ksai_acc, np.logaddexp(ks1, hmm2.ksai_acc), rtol=1e-13, equal_nan=True)
    assert np.allclose(h3.gamma_acc, np.logaddexp(ga1, hmm2.gamma_acc), rtol=1e-13)


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree (build container only)")
def test_accumulator_interchange_with_reference(tmp_path):
    """Our accumulator files through the reference's init_acc + update_param, and the reference's files
    through ours: the same re-estimated parameters."""
    from oracle import ref_harness as rh

    H = rh.Harness(["a"], 4)
    gmm, hmm = _models()
    _fill_acc(gmm, hmm, 5)
    gmm.save_acc(str(tmp_path)); hmm.save_acc(str(tmp_path))
    # ours -> reference
    rg = H.Clustering.GMM(H.log, dimension=39, mix_level=4, gmm_id=2)
    rg.mean, rg.covariance, rg.alpha = gmm.mean.copy(), gmm.covariance.copy(), gmm.alpha.copy()
    rg.init_acc(str(tmp_path))
    assert np.allclose(rg.acc, gmm.acc, rtol=1e-13)
    assert np.allclose(rg.mean_acc, gmm.mean_acc, rtol=1e-13)
    rg.update_param(c_covariance=1e-6)
    mu_ref = gmm._sx / gmm._occ[:, None]
    assert np.allclose(rg.mean, mu_ref, rtol=1e-9, atol=1e-9)
    assert np.allclose(rg.alpha, gmm._occ / gmm._socc, rtol=1e-12)
    assert np.allclose(np.diagonal(rg.covariance, axis1=1, axis2=2), gmm._scc / gmm._occ[:, None], rtol=1e-10)
    ru = H.am.init_unit("a", new_log=True)
    ru.init_acc(str(tmp_path))
    with np.errstate(invalid="ignore"):
        assert np.allclose(ru.ksai_acc, hmm.ksai_acc, equal_nan=True)
    assert np.allclose(ru.gamma_acc, hmm.gamma_acc)
    # reference -> ours
    out = tmp_path / "ref"
    os.mkdir(out)
    rg2 = H.Clustering.GMM(H.log, dimension=39, mix_level=4, gmm_id=2)
    rg2.mean, rg2.covariance, rg2.alpha = gmm.mean.copy(), gmm.covariance.copy(), gmm.alpha.copy()
    rg2.init_acc(str(tmp_path))  # now holds our statistics in the reference's log domain
    rg2.save_acc(str(out))
    ru.save_acc(str(out))
    g4, h4 = _models()
    g4.init_acc(str(out)); h4.init_acc(str(out))
    assert np.allclose(g4._occ, gmm._occ, rtol=1e-12) and np.allclose(g4._sx, gmm._sx, rtol=1e-9, atol=1e-9)
    assert np.allclose(g4._scc, gmm._scc, rtol=1e-12) and np.isclose(g4._socc, gmm._socc, rtol=1e-12)
    with np.errstate(invalid="ignore"):
        assert np.allclose(h4.ksai_acc, hmm.ksai_acc, equal_nan=True)
