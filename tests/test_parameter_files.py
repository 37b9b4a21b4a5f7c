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
