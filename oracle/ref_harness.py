"""Harness that imports and drives the UNMODIFIED reference from /root/reference.

TEST INFRASTRUCTURE ONLY (see oracle/ref_port.py header).  It exists in this container only:
the GPU box has no /root/reference, so nothing marked ``gpu``, ``smoke()`` or ``bench.py`` may
use it.  Consumers: ``tests/golden/make_golden.py`` (fixture generation) and the optional
``tests/test_oracle_vs_reference.py`` (skipped when the reference tree is absent).

Recipe = SURVEY.md Appendix B: env vars before import, pyaudio/pylab stubbed, ``time`` patched
to a counter so accumulator file names never collide (Q13).
"""
from __future__ import annotations

import os
import sys
import tempfile
import types
from unittest.mock import MagicMock

import numpy as np

REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "StatisticalModel"))


class _Quiet:
    def note(self, *a, **k):
        pass

    def close(self):
        pass

    def append(self):
        pass

    def generate(self):
        pass


class Harness:
    """One reference AcousticModel over a temp parameter tree."""

    def __init__(self, units, mix_level, state_num=5, unit_type="T"):
        assert available(), "reference tree not present"
        self.tmp = tempfile.mkdtemp(prefix="poccala_ref_")
        unit_dir = os.path.join(self.tmp, "Unit")
        os.makedirs(os.path.join(unit_dir, "Parameters", unit_type))
        with open(os.path.join(unit_dir, unit_type), "w") as f:
            f.write("title\n" + ",".join(units) + "\n")
        os.environ["unit_file_path"] = unit_dir
        os.environ["parameters_file_path"] = os.path.join(unit_dir, "Parameters")
        os.environ["log_file_path"] = os.path.join(unit_dir, "Parameters")
        sys.modules.setdefault("pyaudio", MagicMock())
        sys.modules.setdefault("pylab", MagicMock())
        sys.dont_write_bytecode = True
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            # a previous Harness may have imported these with other env paths: reload
            for m in ("AcousticModel.AcousticModel", "LogPrint"):
                sys.modules.pop(m, None)
            from AcousticModel.AcousticModel import AcousticModel  # noqa
            from StatisticalModel.LHMM import LHMM  # noqa
            from StatisticalModel.Clustering import Clustering  # noqa
            from LogPrint import Log  # noqa
            import StatisticalModel.LHMM as lhmm_mod
            import StatisticalModel.Clustering as cl_mod
        counter = {"t": 1_000_000}

        def fake_time():
            counter["t"] += 1
            return counter["t"]

        lhmm_mod.time = types.SimpleNamespace(time=fake_time)
        cl_mod.time = types.SimpleNamespace(time=fake_time)
        self.LHMM = LHMM
        self.Clustering = Clustering
        self.Log = Log
        self.unit_type = unit_type
        self.state_num = state_num
        self.mix = mix_level
        self.log = Log(unit_type, console=False)
        self.log.generate()
        self.am = AcousticModel(self.log, unit_type, processes=1, console=False, state_num=state_num,
                                mix_level=mix_level)

    # -- parameters: dict unit -> list of (mean[M,D], var[M,D], alpha[M]) per emitting state, + transmat
    def make_unit(self, name, params, transmat=None):
        h = self.am.init_unit(name, new_log=True)
        for g, (mu, var, al) in zip(h.profunction[1:-1], params):
            g.mean = np.array(mu, dtype=np.float64)
            g.covariance = np.stack([np.diag(v) for v in np.asarray(var, dtype=np.float64)])
            g.alpha = np.array(al, dtype=np.float64)
        if transmat is not None:
            h.change_A(np.array(transmat, dtype=np.float64))
        return h

    def estep_in_memory(self, label, X, params, transmats=None, fix_code=0):
        """Appendix B step 6: the reference's own scoring, sentence HMM, Baum-Welch and update_acc."""
        X = np.asarray(X, dtype=np.float64)
        T = len(X)
        hmm_list = []
        for u in label:
            h = self.make_unit(u, params[u], None if transmats is None else transmats[u])
            h.cal_observation_pro([X], [T])
            h.clear_data()
            hmm_list.append(h)
        states, A, B, pi = self.am.embedded(label, hmm_list, 0, 15)
        eh = self.LHMM(states, self.state_num, _Quiet(), transmat=A, probmat=[B], pi=pi, hmm_list=hmm_list,
                       fix_code=fix_code)
        eh.add_data([X])
        eh.add_T([T])
        eh.baulm_welch()
        return dict(states=states, A=A, B=B, pi0=pi, eh=eh, hmm_list=hmm_list,
                    alpha=eh._LHMM__result_f[0], beta=eh._LHMM__result_b[0],
                    ksai=eh._LHMM__ksai, gamma=eh._LHMM__gamma, pi_next=eh.pi)

    def viterbi(self, states, A, B, pi):
        return self.LHMM.viterbi(_Quiet(), states, A, B, pi, convert=False)

    def viterbi_labels(self, states, A, B, pi):
        self.am.log = _Quiet()
        return self.am.viterbi(states, A, B, pi)

    def save_params(self, name, h):
        self.am._AcousticModel__save_parameter(name, h)

    def file_based_iteration(self, units_params, utterances, c_cov, transmats=None):
        """Appendix B step 7: the reference's exact file-based code path for one EM iteration."""
        for name, p in units_params.items():
            h = self.make_unit(name, p, None if transmats is None else transmats[name])
            self.save_params(name, h)
        n = len(utterances)
        for i, (label, X) in enumerate(utterances):
            self.am.log = self.Log(self.unit_type, console=False)
            self.am.log.append()
            self.am.multi_embedded_training_1(list(label), np.asarray(X, dtype=np.float64), False, False, i + 1, n, 0)
        out = {}
        names = list(units_params.keys())
        for k, name in enumerate(names):
            self.am.log = self.Log(self.unit_type, console=False)
            self.am.log.append()
            self.am.multi_embedded_training_2(name, False, False, False, c_cov, k + 1, len(names), len(names), 0)
            h = self.am.init_unit(name, new_log=False)
            self.am.init_parameter(name, h)
            out[name] = dict(
                transmat=np.array(h.transmat),
                gmms=[dict(mean=np.array(g.mean), var=np.array([np.diag(c) for c in g.covariance]),
                           alpha=np.array(g.alpha)) for g in h.profunction[1:-1]],
            )
        return out

    def kmeans(self, data, k, seed, cov_matrix=True):
        import random

        random.seed(seed)
        ci = self.Clustering.ClusterInitialization(data, k, len(data[0]), _Quiet())
        return ci.kmeans(algorithm=1, cov_matrix=cov_matrix)
