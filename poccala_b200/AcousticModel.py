"""Device-backed mirror of the reference's AcousticModel/AcousticModel.py training-loop surface for
the embedded Baum-Welch path (AcousticModel.py:842-1043): same class name, constructor arguments
and method names; utterances live in memory / HBM instead of pickled files, the accumulator
fan-in is a device reduction (+ one NCCL allreduce when a process group is given) instead of
timestamp-named .npy files (LHMM.py:211-290, Clustering.py:257-367).

Two ways through the same kernels:
  * batched (the product path): `add_corpus(labels, data)` once, then `embedded_training(...)` per
    EM iteration - K1, K2, K3, the transition reduction and the M-step run over the whole resident
    corpus in a handful of launches;
  * per utterance, object by object, exactly in the reference's call order:
    `multi_embedded_training_1(label, data, init, show_q, i, n, fix_code)` then
    `multi_embedded_training_2(unit, init, ...)` - used by the parity tests that replay the
    reference's own sequence of calls.
Out of scope here: wav/MFCC front end, pickled data files, trainInfo resume files,
multiprocessing pools.
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine as _eng
from . import synth as _synth
from .Clustering import Clustering
from .LHMM import LHMM
from .runtime import NullLog, get_engine

EMIT = _eng.EMIT


class ModeError(Exception):
    """Exceptions.ModeError: a training mode other than 1 / 2 (AcousticModel.py:732, 800)."""


class UnitFileExistsError(Exception):
    """Exceptions.UnitFileExistsError: the unit inventory file is missing (AcousticModel.py:143-144)."""


class AcousticModel(object):
    def __init__(self, log=None, unit_type="IF", mode=0, processes=None, job_id=0, console=True, state_num=5,
                 mix_level=1, dct_num=13, delta_1=True, delta_2=True, unit_file_path=None, device=None):
        """AcousticModel.py:32-107.  `unit_file_path`: directory holding the unit inventory files
        (the reference reads it from the environment, AcousticModel.py:25-26); defaults to the IF
        inventory shipped with this package."""
        self.log = log if log is not None else NullLog()
        self.__unit_type = unit_type
        self.__mode = mode
        self.processes = processes
        self.__job_id = job_id
        self.__console = console
        self.__state_num = state_num
        self.__mix_level = mix_level
        self.__vector_size = dct_num * (1 + int(bool(delta_1)) + int(bool(delta_2)))
        self.__dct_num, self.__delta_1, self.__delta_2 = dct_num, delta_1, delta_2
        self.__unit_file_path = unit_file_path
        self.__device = device
        self.__loaded_units = []
        self.__params = {}   # unit -> dict(mean[3,M,D], var[3,M,D], alpha[3,M], transmat[5,5])
        self.__acc = {}      # unit -> dict(ksai, gamma, gmm=[(occ, socc, sx, scc)]*3) for the per-utterance path
        self.__corpus = None
        self.__frames = None
        self.__utterances = None
        self.__state_data = None  # per-(unit, state) data sets grouped by process_data(mode=1)
        self.__estep = None
        self.__estep_group = None  # process group the resident frames were standardised with
        self.__model = None
        self.__unit_data = {}      # unit -> list of [n, D] segments (multi_process_data: the reference's pickles)
        self.last_log_likelihood = None
        if state_num != 5:
            raise NotImplementedError("the kernels cover state_num = 5 (3 emitting states, init.py:33)")

    # ---- inventory --------------------------------------------------------------------------
    @property
    def loaded_units(self):
        return self.__loaded_units

    @property
    def statenum(self):
        return self.__state_num

    @property
    def engine(self):
        return get_engine(self.__device)

    def load_unit(self, unit_type=None):
        """AcousticModel.py:134-162: first line of the file is a title, the rest comma-separated."""
        import os

        if unit_type:
            self.__unit_type = unit_type
        if self.__unit_file_path is None:
            path = os.path.join(os.path.dirname(_synth.reference_unit_file()), self.__unit_type)
        else:
            path = os.path.join(self.__unit_file_path, self.__unit_type)
        if not os.path.exists(path):
            raise UnitFileExistsError(self.__unit_type)
        self.__loaded_units = _synth.load_unit_file(path)
        self.log.note("units loaded: %d" % len(self.__loaded_units), cls="i")
        return self.__loaded_units

    def set_units(self, units):
        self.__loaded_units = list(units)

    # ---- parameters (in memory; the reference keeps them in .npy files, AcousticModel.py:228-265) ----
    def set_parameters(self, mean, var, alpha, transmat=None):
        """mean/var [U,3,M,D], alpha [U,3,M], transmat [U,5,5] in `loaded_units` order."""
        U = len(self.__loaded_units)
        tm = _synth.default_transmat(U) if transmat is None else np.asarray(transmat, dtype=np.float64)
        for i, u in enumerate(self.__loaded_units):
            self.__params[u] = dict(mean=np.array(mean[i], dtype=np.float64), var=np.array(var[i], dtype=np.float64),
                                    alpha=np.array(alpha[i], dtype=np.float64), transmat=np.array(tm[i]))
        self.__model = None

    def get_parameters(self):
        us = self.__loaded_units
        return (np.stack([self.__params[u]["mean"] for u in us]), np.stack([self.__params[u]["var"] for u in us]),
                np.stack([self.__params[u]["alpha"] for u in us]), np.stack([self.__params[u]["transmat"] for u in us]))

    def flat_start(self, data, proportion=0.25, step=1, differentiation=True, coefficient=1.):
        """AcousticModel.py:479-517 (`__flat_start`) on in-memory utterances: the frames of the first
        int(len(data) * proportion) utterances, every `step`-th frame, go through
        ClusterInitialization(k=1).kmeans(algorithm=1, cov_matrix=True) - i.e. the global mean and
        variance, by the device k-means with the reference's insertion-order sums - and every GMM
        of every unit starts from mean + diff * variance (diff drawn from numpy's global RNG exactly
        like the reference, one [mix, 1] draw shared by all units) and the global variance."""
        n = int(len(data) * proportion)
        p_data = np.asarray(data[0], dtype=np.float64)[::step]
        for index in range(1, n):
            p_data = np.append(p_data, np.asarray(data[index], dtype=np.float64)[::step], axis=0)
        cluster = Clustering.ClusterInitialization(p_data, 1, self.__vector_size, self.log)
        mean, covariance, alpha, clustered = cluster.kmeans(algorithm=1, cov_matrix=True)
        covariance_diagonal = covariance[0].diagonal()
        M = self.__mix_level
        diff = np.zeros((M, 1))
        if differentiation:
            assert 0 <= coefficient <= 1, "coefficient must lie in [0, 1]"
            diff = (np.random.random((M, 1)) - np.random.random((M, 1))) * coefficient
        g_mean = mean.repeat(M, axis=0) + diff * covariance_diagonal
        g_var = np.repeat(covariance_diagonal[None], M, axis=0)
        tm = _synth.default_transmat(1)[0]
        for unit in self.__loaded_units:
            self.__params[unit] = dict(mean=np.stack([g_mean] * EMIT), var=np.stack([g_var] * EMIT),
                                       alpha=np.ones((EMIT, M)) / M, transmat=tm.copy())
        self.__model = None
        self.delete_trainInfo()

    def init_unit(self, unit=None, new_log=True, fix_code=0):
        """AcousticModel.py:164-226: a 5-state left-to-right unit HMM: entry VirtualState(1.), three
        GMM states, exit VirtualState(0.); transmat [0,1] = 1, emitting rows (0.5, 0.5)."""
        D, M = self.__vector_size, self.__mix_level
        gmms = [Clustering.GMM(self.log, dimension=D, mix_level=M, gmm_id=i) for i in range(1, self.__state_num - 1)]
        profunc = [AcousticModel.VirtualState(1.)] + gmms + [AcousticModel.VirtualState(0.)]
        states = {i: unit for i in range(self.__state_num)}
        hmm = LHMM(states, self.__state_num, self.log, transmat=_synth.default_transmat(1)[0], profunc=profunc,
                   fix_code=fix_code)
        return hmm

    def init_parameter(self, unit, hmm):
        """AcousticModel.py:228-240: load the unit's current parameters into `hmm`."""
        p = self.__params.get(unit)
        if p is None:
            return
        hmm.change_A(p["transmat"].copy())
        for r, g in enumerate(hmm.profunction[1:-1]):
            g.mean = p["mean"][r].copy()
            g.covariance = np.stack([np.diag(v) for v in p["var"][r]])
            g.alpha = p["alpha"][r].copy()

    def __save_parameter(self, unit, hmm):
        self.__params[unit] = dict(
            mean=np.stack([g.mean for g in hmm.profunction[1:-1]]),
            var=np.stack([g.variance for g in hmm.profunction[1:-1]]),
            alpha=np.stack([g.alpha for g in hmm.profunction[1:-1]]),
            transmat=np.array(hmm.transmat, dtype=np.float64))
        self.__model = None

    def __save_acc(self, unit, hmm):
        """The reference writes one .npy per accumulator per call (AcousticModel.py:267-287) and
        merges them with a log-sum-exp over the files; here the merge happens immediately."""
        a = self.__acc.get(unit)
        gm = [(g._occ.copy(), g._socc, g._sx.copy(), g._scc.copy()) for g in hmm.profunction[1:-1]]
        if a is None:
            self.__acc[unit] = dict(ksai=np.array(hmm.ksai_acc), gamma=np.array(hmm.gamma_acc), gmm=gm)
        else:
            a["ksai"] = np.logaddexp(a["ksai"], hmm.ksai_acc)
            a["gamma"] = np.logaddexp(a["gamma"], hmm.gamma_acc)
            a["gmm"] = [(o0 + o1, s0 + s1, x0 + x1, c0 + c1) for (o0, s0, x0, c0), (o1, s1, x1, c1) in zip(a["gmm"], gm)]

    def delete_buffer_file(self, unit, show_info=False):
        """AcousticModel.py:376-430 deletes the accumulator files; here: drop the in-memory ones."""
        self.__acc.pop(unit, None)

    def delete_trainInfo(self):
        """AcousticModel.py:432-441 removes trainInfo_<job>.csv (the resume list of trained units); training
        state lives in memory here, so there is nothing to delete."""

    @staticmethod
    def init_audio(audiopath, labelpath):
        """AcousticModel.py:443-461: generator over a corpus directory - first the number of audio files,
        then (wav path, transcript path) pairs, each terminated by a newline like the reference's
        pathInfo lines."""
        import os

        count = 0
        for d in os.walk(audiopath):
            count += len(d[2])
        yield count
        for d in os.walk(audiopath):
            for file in d[2]:
                name = file.split('.')[0]
                yield audiopath + '/%s.wav\n' % name, labelpath + '/%s.wav.trn\n' % name

    def load_audio(self, audiopath):
        """AcousticModel.__load_audio (AcousticModel.py:463-477): MFCC features of a wave file (with the first and
        second differences the model was configured with), frames the cepstral-distance detector rejects removed.
        Computed on the device (csrc/mfcc.cu)."""
        from .AudioProcessing import AudioProcessing

        mfcc = AudioProcessing.MFCC(self.__dct_num)
        mfcc.init_audio(path=audiopath)
        m = mfcc.mfcc(nfft=512, d1=self.__delta_1, d2=self.__delta_2)
        vad = AudioProcessing.VAD()
        vad.init_mfcc(m)
        return vad.mfcc()

    # ---- sentence HMM assembly (AcousticModel.py:957-1014) ----------------------------------
    def embedded(self, label, hmm_list, data_index, alter=15):
        """states {index: unit}, A [N,N] (unit rows pasted so that a unit's exit column is the next
        unit's first emitting state), B [N,T] (entry row of the first unit, emitting rows of all,
        exit row of the last), pi uniform 1/N over ALL states (Q4); `alter` selects which to build."""
        S = self.__state_num
        E = S - 2
        N = E * len(hmm_list) + 2
        out = []
        if alter & 8:
            names = [label[0]] + [u for u in label for _ in range(E)] + [label[-1]]
            out.append(dict(enumerate(names)))
        if alter & 4:
            A = np.zeros((N, N))
            A[:S - 1, :S] = hmm_list[0].transmat[:-1]
            for i, h in enumerate(hmm_list):
                a = i * E + 1
                A[a:a + E, a - 1:a - 1 + S] = h.transmat[1:-1]
            out.append(A)
        if alter & 2:
            rows = [hmm_list[0].B_p[data_index][0:1]]
            rows += [h.B_p[data_index][1:-1] for h in hmm_list]
            rows.append(hmm_list[-1].B_p[data_index][-1:])
            out.append(np.concatenate(rows, axis=0))
        if alter & 1:
            out.append(np.ones((N,)) / N)
        return out

    def viterbi(self, complex_states, complex_transmat, complex_prob, complex_pi):
        """AcousticModel.py:1016-1027: forced alignment, per-frame unit labels."""
        return LHMM.viterbi(self.log, complex_states, complex_transmat, complex_prob, complex_pi, convert=True,
                            show_mark_state=True)

    @staticmethod
    def discriminate(unit, sequence):
        """AcousticModel.py:937-955: split the frames labelled `unit` into contiguous runs (index
        bookkeeping on the alignment the device produced)."""
        loc = np.where(np.asarray(sequence) == unit)[0]
        if len(loc) == 0:
            return []
        cuts = np.where(np.diff(loc) != 1)[0] + 1
        return np.split(loc, cuts)

    # ---- per-utterance path, in the reference's call order -----------------------------------
    def multi_embedded_training_1(self, label, data, init, *args):
        """AcousticModel.py:884-916.  args = (show_q, current, total, fix_code)."""
        show_q = args[0] if len(args) > 0 else False
        fix_code = args[3] if len(args) > 3 else 0
        data = np.asarray(data, dtype=np.float64)
        hmm_list = []
        for unit in label:
            hmm = self.init_unit(unit=unit, new_log=init)
            self.init_parameter(unit, hmm=hmm)
            hmm_list.append(hmm)
            hmm.cal_observation_pro([data], [len(data)])
            hmm.clear_data()
        states, A, B, pi = self.embedded(label, hmm_list, 0, 15)
        embed_hmm = LHMM(states, self.__state_num, self.log, transmat=A, probmat=[B], pi=pi, hmm_list=hmm_list,
                         fix_code=fix_code)
        embed_hmm.add_data([data])
        embed_hmm.add_T([len(data)])
        embed_hmm.baulm_welch(show_q=show_q)
        for i in range(len(label)):
            self.__save_acc(label[i], hmm_list[i])
        return embed_hmm

    def multi_embedded_training_2(self, unit, init, *args):
        """AcousticModel.py:918-935.  args = (show_q, show_a, c_covariance, done, todo, total, fix_code)."""
        show_q = args[0] if len(args) > 0 else False
        show_a = args[1] if len(args) > 1 else False
        c_cov = args[2] if len(args) > 2 else 1e-3
        fix_code = args[6] if len(args) > 6 else 0
        hmm = self.init_unit(unit, new_log=init, fix_code=fix_code)
        self.init_parameter(unit, hmm)
        a = self.__acc.get(unit)
        if a is not None:
            hmm.add_acc(a["ksai"], a["gamma"])
            for g, (occ, socc, sx, scc) in zip(hmm.profunction[1:-1], a["gmm"]):
                g._occ, g._socc, g._sx, g._scc = occ.copy(), socc, sx.copy(), scc.copy()
            hmm.update_param(show_q=show_q, show_a=show_a, c_covariance=c_cov)
        self.__save_parameter(unit, hmm)
        return hmm

    # ---- batched path ---------------------------------------------------------------------------
    def add_corpus(self, labels, data):
        """labels: list of unit-name lists, data: list of [T,D] arrays (or one [sum T, D] device
        tensor with `labels` and frame counts given as (labels, n_frames)).  Frames are copied to
        HBM once and stay resident across EM iterations."""
        eng = self.engine
        idx = {u: i for i, u in enumerate(self.__loaded_units)}
        lab = [np.array([idx[u] for u in l], dtype=np.int32) for l in labels]
        n_frames = np.array([len(x) for x in data], dtype=np.int32)
        self.__corpus = _eng.Corpus(eng, lab, n_frames, len(self.__loaded_units))
        self.__frames = torch.as_tensor(np.concatenate([np.asarray(x) for x in data], axis=0)).to(eng.device)
        self.__utterances = data
        self.__state_data = None
        self.__estep = None
        return self.__corpus

    def _ensure_estep(self, group=None):
        if self.__corpus is None:
            raise RuntimeError("add_corpus(labels, data) first")
        if self.__model is None:
            self.__model = _eng.Model(self.engine, *self.get_parameters())
            self.__estep = None
        if self.__estep is not None and group is not None and self.__estep_group is not group:
            # the frames were standardised with another group's (or this rank's own) moments: the accumulators of
            # the ranks would live in different coordinates - standardise again with the moments of `group`
            self.__estep = None
        if self.__estep is None:
            self.__estep = _eng.EStep(self.engine, self.__corpus, self.__model)
            self.__estep.load_frames(self.__frames, group=group)
            self.__estep_group = group
        return self.__estep

    def embedded_training(self, wwt_units=None, init=True, load_line=0, fix_code=0, show_q=False, show_a=False,
                          c_covariance=1e-3, group=None):
        """AcousticModel.py:842-882: one embedded Baum-Welch iteration over the resident corpus.
        `group`: torch.distributed process group - every rank holds its own shard of the
        utterances and the accumulators are allreduced (the reference merges accumulator files
        from all machines, LHMM.py:256-290)."""
        es = self._ensure_estep(group)
        es.em_iteration(c_covariance=c_covariance, fix_code=fix_code, group=group)
        mean, var, alpha, tm = self.__model.numpy()
        keep = None if wwt_units is None else set(wwt_units)
        for i, u in enumerate(self.__loaded_units):
            if keep is not None and u not in keep:
                continue  # only the requested units take their new parameters (AcousticModel.py:872-880)
            self.__params[u] = dict(mean=mean[i], var=var[i], alpha=alpha[i], transmat=tm[i])
        if keep is not None and len(keep) != len(self.__loaded_units):
            self.__model = None  # rebuild from the selectively updated parameters
        self.last_log_likelihood = float(es.utt_logp.sum())
        if show_q:
            self.log.note("sum log P(O) = %f" % self.last_log_likelihood, cls="i")
        return self.last_log_likelihood

    # ---- mode 1: segment -> per-state data sets -> k-means / GMM EM --------
    def process_data(self, mode=1, load_line=0, init=True, proportion=0.25, step=1, differentiation=True,
                     coefficient=1, group=None):
        """AcousticModel.py:681-733 over the resident corpus.  Mode 1: every utterance is cut
        uniformly over its units (init, `__eq_segment` mode 'e', :605-612) or along the forced
        alignment of the current models (multi_process_data, :736-764), each unit segment in three
        (`__get_gmmdata`, :630-644), and the frames are grouped per (unit, state) on the device -
        what the reference spreads over one pickle per segment.  Mode 2 with init: the flat start."""
        if self.__corpus is None:
            raise RuntimeError("add_corpus(labels, data) first")
        if mode == 2:
            if init:
                self.flat_start(self.__utterances, proportion=proportion, step=step, differentiation=differentiation,
                                coefficient=coefficient)
            return None
        if mode != 1:
            raise ModeError("mode must be 1 or 2 (Exceptions.ModeError)")
        path = None
        if not init:
            self.log.note("Viterbi Alignment...", cls="i")
            _, path, _ = self.align(group)
        key, kept = _eng.segment_keys(self.engine, self.__corpus, path)
        frames = self.__frames if self.__frames.dtype == torch.float64 else self.__frames.double()
        self.__state_data = _eng.group_frames(self.engine, key, EMIT * len(self.__loaded_units), frames)
        self.__state_data["utt_kept"] = kept
        return self.__state_data

    def multi_process_data(self, label, data, init, *args):
        """AcousticModel.py:723-769, one utterance: cut `data` uniformly over the units of `label` (init,
        `__eq_segment` mode 'e') or along the forced alignment of the current models (Viterbi on the
        device; an utterance whose path visits fewer distinct units than its label holds is dropped,
        :751-757; contiguous runs per unit, `discriminate`), and keep the segments per unit - what the
        reference pickles with `__save_data`.  `multi_training` reads them when no corpus-wide grouping
        (process_data) is at hand.  args = (current, total, fix_code), unused."""
        data = np.asarray(data, dtype=np.float64)
        label = list(label)
        if init:
            n = len(data) // len(label)
            for i, u in enumerate(label):
                self.__unit_data.setdefault(u, []).append(data[i * n:(i + 1) * n])
            return True
        hmm_list = []
        for u in label:
            hmm = self.init_unit(unit=u, new_log=init)
            self.init_parameter(u, hmm=hmm)
            hmm.cal_observation_pro([data], [len(data)])
            hmm.clear_data()
            hmm_list.append(hmm)
        states, A, B, pi = self.embedded(label, hmm_list, 0, 15)
        point, sequence = self.viterbi(states, A, B, pi)
        if len(set(sequence)) < len(set(label)):
            self.log.note("viterbi alignment failed: utterance dropped", cls="w")
            return False
        for u in set(label):
            for loc in AcousticModel.discriminate(u, sequence):
                self.__unit_data.setdefault(u, []).append(data[loc])
        return True

    def _state_rows_from_segments(self, unit):
        """`__get_gmmdata` (AcousticModel.py:629-644) on the segments multi_process_data kept: every segment of n
        frames is cut into 3 parts of n // 3 frames, the last taking the remainder; parts concatenate per state."""
        segs = self.__unit_data.get(unit, [])
        D = self.__vector_size
        parts = [[] for _ in range(EMIT)]
        for seg in segs:
            n = len(seg) // EMIT
            for r in range(EMIT):
                parts[r].append(seg[r * n:(r + 1) * n] if r < EMIT - 1 else seg[r * n:])
        dev = self.engine.device
        return [torch.as_tensor(np.concatenate(p) if p else np.zeros((0, D))).to(dev) for p in parts]

    def multi_training(self, unit, init, *args):
        """AcousticModel.py:814-838 + `__cal_gmm` (:532-561): the unit's three state GMMs from their
        data sets: k-means initialisation (ClusterInitialization.kmeans(algorithm=1), all three
        states in one launch) when `init` or when the mixture count changed, then GMM.em.  A state
        with fewer frames than mixtures is skipped (:546-548); a unit without data keeps its
        parameters (:830-832).  args = (show_q, show_a, c_covariance, ...).  As in the reference the EM
        runs with the split / merge search when `init` is set (em(smem=init), :835 -> :560)."""
        import random

        show_q = args[0] if len(args) > 0 else False
        c_cov = args[2] if len(args) > 2 else 1e-3
        if self.__state_data is None and not self.__unit_data:
            raise RuntimeError("process_data(mode=1) or multi_process_data(...) first")
        hmm = self.init_unit(unit, new_log=init, fix_code=2)
        self.init_parameter(unit, hmm)
        M = self.__mix_level
        if self.__state_data is not None:
            ui = self.__loaded_units.index(unit)
            off, data = self.__state_data["key_off"], self.__state_data["data"]
            rows = [data[off[ui * EMIT + r]:off[ui * EMIT + r + 1]] for r in range(EMIT)]
        else:
            rows = self._state_rows_from_segments(unit)
        if sum(len(x) for x in rows) == 0:
            self.log.note("unit %s has no data" % unit, cls="w")
            self.__save_parameter(unit, hmm)
            return hmm
        # state by state, as __cal_gmm does (AcousticModel.py:547-561): the seeding draws, the one draw per k-means
        # move (Clustering.py:932) and the draws of the split / merge search all come from the same `random` stream, so
        # the order of the states is part of the result
        for r in range(EMIT):
            gmm = hmm.profunction[r + 1]
            if len(rows[r]) < M:
                gmm.log.note("too little data, state skipped", cls="w")
                continue
            if init or gmm.mixture != M:
                seeds = _eng.kmeans_seed_points(np.ascontiguousarray(rows[r][:, 0].cpu().numpy()), M, random)
                km = _eng.kmeans_run(self.engine, rows[r], np.array([0, len(rows[r])], dtype=np.int64), M,
                                     np.array([seeds], dtype=np.int32))
                for _ in range(int(km["moves"][0])):
                    random.random()
                gmm.mixture = M
                gmm.mean = km["mean"][0].cpu().numpy()
                gmm.covariance = np.stack([np.diag(v) for v in km["var"][0].cpu().numpy()])
                gmm.alpha = km["alpha"][0].cpu().numpy()
            gmm.add_data(rows[r])
            gmm.em(show_q=show_q, smem=init, c_covariance=c_cov)
            gmm.clear_data()
        self.__save_parameter(unit, hmm)
        return hmm

    def training(self, mode=2, init=True, show_q=False, show_a=False, load_line=0, c_covariance=1e-3, group=None):
        """AcousticModel.py:771-813.  Mode 1: per-unit GMM training on the data sets process_data
        grouped, then one embedded iteration that re-estimates the transitions only (fix_code 2);
        mode 2: one embedded iteration of everything."""
        units = list(self.__loaded_units)
        if mode == 1:
            fix_code = 2
            for n, unit in enumerate(units):
                self.multi_training(unit, init, show_q, show_a, c_covariance, n + 1, len(units), len(units))
        elif mode == 2:
            fix_code = 0
        else:
            raise ModeError("mode must be 1 or 2 (Exceptions.ModeError)")
        return self.embedded_training(units, init=init, show_q=show_q, show_a=show_a, load_line=load_line,
                                      fix_code=fix_code, c_covariance=c_covariance, group=group)

    def align(self, group=None):
        """Forced alignment of the whole resident corpus (multi_process_data's Viterbi step,
        AcousticModel.py:736-764): returns (scores [U] fp64, state path, unit path) as device tensors
        indexed by corpus frame.  `group`: the process group of a data-parallel run, so that the frames
        are standardised once, with the moments of all ranks."""
        es = self._ensure_estep(group)
        es.score()
        ls, ln = _eng.host_log_bands(self.get_parameters()[3], self.engine.device)
        n_lab = self.__corpus.n_labels
        with np.errstate(divide="ignore"):
            logpi = np.array([np.log(np.ones(EMIT * l + 2) / (EMIT * l + 2))[0] for l in n_lab])
        return _eng.viterbi(self.engine, self.__corpus, es.b, ls, ln,
                            utt_logpi=torch.as_tensor(logpi).to(self.engine.device))

    class VirtualState(object):
        """AcousticModel.py:1029-1043: scoring stub of the non-emitting entry / exit states."""

        def __init__(self, p=0.):
            self.__p = p

        def point(self, x, log=False, standard=False, record=False):
            if log:
                with np.errstate(divide="ignore"):
                    return np.log(self.__p)
            return self.__p
