"""Device-backed mirror of the reference's StatisticalModel/LHMM.py `LHMM` class for the embedded
Baum-Welch hot path: same constructor, properties and method names (LHMM.py:19-609), numbers from
the sm_100a kernels behind include/poccala_b200.h (K1 scoring, K2 forward-backward, K3
accumulation, K4 Viterbi).  There is no CPU path.

Scope: the only way Baum-Welch works in the reference is on the sentence HMM that
AcousticModel.embedded assembles (left-to-right, bidiagonal, entry state emitting log 1, exit state
log 0, uniform pi) with `probmat=[B]` and `hmm_list`; that is the structure the kernels cover.
Anything else (dense transition matrices, profunc-only models: Q16 crashes in the reference too)
raises `UnsupportedModel`.  File IO (save_parameter / save_acc / init_parameter / init_acc,
LHMM.py:192-290) is the on-disk format row (§8 f1) and is not part of this path.
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine as _eng
from .Clustering import Clustering
from .runtime import NullLog, get_engine

EMIT = _eng.EMIT


class UnsupportedModel(NotImplementedError):
    pass


def _lse(a, axis=None):
    a = np.asarray(a, dtype=np.float64)
    m = np.max(a, axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0.0)
    with np.errstate(divide="ignore"):
        out = m + np.log(np.sum(np.exp(a - m), axis=axis, keepdims=True))
    return out if axis is None else np.squeeze(out, axis=axis)


def _banded_units(transmat, n_emit=EMIT):
    """Split a composite [N,N] left-to-right transition matrix (AcousticModel.embedded,
    AcousticModel.py:978-988) into one pseudo unit matrix [5,5] per label position; raises when
    the matrix has mass outside the (self, next) band."""
    A = np.asarray(transmat, dtype=np.float64)
    N = A.shape[0]
    if A.shape != (N, N) or (N - 2) % n_emit or N < n_emit + 2:
        raise UnsupportedModel("composite transmat must be [3L+2, 3L+2]")
    band = np.zeros_like(A)
    idx = np.arange(N)
    band[idx, idx] = A[idx, idx]
    band[idx[:-1], idx[:-1] + 1] = A[idx[:-1], idx[:-1] + 1]
    if np.any(band != A):
        raise UnsupportedModel("transition mass outside the (self, next) band: not a left-to-right sentence HMM")
    L = (N - 2) // n_emit
    tm = np.zeros((L, 5, 5))
    for p in range(L):
        a = 1 + n_emit * p
        tm[p, 1:1 + n_emit, :] = A[a:a + n_emit, a - 1:a - 1 + 5]
    tm[0, 0, :5] = A[0, :5]
    return tm


class LHMM(object):
    def __init__(self, states, statesnum, log, t=None, transmat=None, profunc=None, probmat=None, pi=None,
                 hmm_list=None, fix_code=0):
        """LHMM.py:19-96 (same arguments).  `states` {index: unit}, `statesnum` states of one unit
        HMM (5), `probmat` = [B[N,T]] emissions that take part in the recurrences, `hmm_list` the
        unit HMMs of the sentence (they receive the accumulators), `fix_code` bits
        (transmat, profunc, pi) = (4, 2, 1)."""
        self.__states = states
        self.__statesnum = statesnum
        self.__hmm_size = len(states)
        self.log = log if log is not None else NullLog()
        self.__t = [] if t is None else t
        n = len(states)
        self.__transmat = np.ones((n, n)) / n if transmat is None else transmat
        self.__pi = np.ones((n,)) / n if pi is None else pi
        assert profunc is not None or probmat is not None, "one of profunc / probmat is required"
        self.__profunction = profunc
        self.__result_p = probmat
        self.__data = []
        with np.errstate(divide="ignore"):
            self.__ksai_acc = np.log(np.zeros((statesnum - 2, statesnum)))
            self.__gamma_acc = np.log(np.zeros((statesnum - 2,)))
        self.__hmm_list = [self] if hmm_list is None else hmm_list
        self.__fix_code = None
        self.__fix_list = None
        self.fix_code = fix_code
        self.log_likelihood = None   # log P(O) of the last baulm_welch
        self.iterations = None       # pi-iterations the reference's loop runs (LHMM.py:526-544)

    # ---- DataInitialization surface (DataInitialization.py:32-120) --------------------------------
    def init_data(self, data=None, datapath=None, shuffle=False, continuous=True, matrix=True, hasmark=True):
        """DataInitialization.py:32-90: take `data` as is, or read the reference's text format - a title
        line, a line `<n> <dimension> <k> <mark 1> .. <mark k>`, then one comma-separated sample per line
        whose last field is its mark."""
        import random

        if data is not None:
            self.__data = data
        elif datapath is not None:
            self.__markdata, self.__data = {}, []
            with open(datapath) as f:
                f.readline()
                s = f.readline().strip('\n').split(' ')
                self.__dimension, self.__classes = int(s[1]), int(s[2])
                if hasmark:
                    for i in range(self.__classes):
                        self.__markdata[s[i + 3]] = []
                for line in f:
                    fields = line.strip('\n').split(',')
                    if len(fields) < self.__dimension:
                        continue
                    row = [float(x) if continuous else x for x in fields[:self.__dimension]]
                    row = np.array(row) if matrix else row
                    if hasmark:
                        self.__markdata[fields[self.__dimension]].append(row)
                    self.__data.append(row)
        else:
            raise Exception("init_data: neither data nor datapath given")
        if shuffle:
            random.shuffle(self.__data)

    def add_data(self, data):
        self.__data.extend(data)

    def clear_data(self):
        self.__data = []

    @property
    def data(self):
        return self.__data

    @property
    def datasize(self):
        return len(self.__data)

    @property
    def markdata(self):
        return getattr(self, "_LHMM__markdata", {})

    @property
    def dimension(self):
        return getattr(self, "_LHMM__dimension", 0)

    @property
    def classes(self):
        return getattr(self, "_LHMM__classes", 0)

    # ---- properties (LHMM.py:100-145) ----------------------------------------------------------
    @property
    def states(self):
        return self.__states

    @property
    def t(self):
        return self.__t

    @property
    def transmat(self):
        return self.__transmat

    @property
    def B_p(self):
        return self.__result_p

    @property
    def pi(self):
        return self.__pi

    @property
    def profunction(self):
        return self.__profunction

    @property
    def ksai_acc(self):
        return self.__ksai_acc

    @property
    def gamma_acc(self):
        return self.__gamma_acc

    @property
    def fix_code(self):
        return self.__fix_code

    @fix_code.setter
    def fix_code(self, fix_code):
        self.__fix_code = fix_code
        self.__fix_list = [bool(fix_code & 2 ** e) for e in range(2, -1, -1)]
        if self.__profunction is None and self in self.__hmm_list:
            self.__fix_list[1] = True

    def change_t(self, t):
        self.__t = t

    def add_T(self, t):
        self.__t.extend(t)

    def change_pi(self, pi):
        self.__pi = pi

    def change_A(self, A):
        self.__transmat = A

    def clear_result_buffer(self):
        self.__result_p = None

    # ---- parameter files (LHMM.py:192-209, 233-254): same layout as the reference ---------------
    def save_parameter(self, path):
        import configparser
        import os

        path = path + '/HMM'
        if not os.path.exists(path):
            os.mkdir(path)
        np.save(path + '/transmat.npy', self.__transmat)
        np.save(path + '/pi.npy', self.__pi)
        with open(path + '/HMM_config.ini', 'w+') as f:
            cfg = configparser.ConfigParser()
            cfg.add_section('Configuration')
            cfg.set('Configuration', 'FIX_CODE', value=str(self.__fix_code))
            cfg.write(f)

    def init_parameter(self, path):
        path = path + '/HMM'
        self.__transmat = np.load(path + '/transmat.npy')
        self.__pi = np.load(path + '/pi.npy')
        # HMM_config.ini is written but never parsed by the reference (Q15): fix_code keeps its value

    # ---- accumulator files (LHMM.py:211-231, 256-290): the reference's layout, so partial E-steps written by
    # either side merge in the other.  The reference names the files by int(time.time()) and loses one of two
    # saves of the same second (Q13); here a counter is appended when the name is taken.
    @staticmethod
    def _acc_file(directory, stem):
        import os
        import time

        name = "%s_%d.npy" % (stem, int(time.time()))
        k = 0
        while os.path.exists(os.path.join(directory, name)):
            k += 1
            name = "%s_%d_%d.npy" % (stem, int(time.time()), k)
        return os.path.join(directory, name)

    def save_acc(self, path):
        import os

        path = path + '/HMM'
        for d in (path, path + '/ksai-acc', path + '/gamma-acc'):
            if not os.path.exists(d):
                os.mkdir(d)
        np.save(self._acc_file(path + '/ksai-acc', 'ksai_acc'), self.__ksai_acc)
        np.save(self._acc_file(path + '/gamma-acc', 'gamma_acc'), self.__gamma_acc)

    def init_acc(self, path):
        """Log-sum-exp of every accumulator file under <path>/HMM (util.matrix_log_sum_exp / log_sum_exp)."""
        import os

        path = path + '/HMM'
        kdir, gdir = path + '/ksai-acc', path + '/gamma-acc'
        if not (os.path.exists(kdir) and os.path.exists(gdir)):
            return
        ks = [np.load(os.path.join(kdir, f)) for f in sorted(os.listdir(kdir))]
        gs = [np.load(os.path.join(gdir, f)) for f in sorted(os.listdir(gdir))]
        if ks:
            self.__ksai_acc = _lse(np.stack(ks), axis=0)
        if gs:
            self.__gamma_acc = _lse(np.stack(gs), axis=0)

    # ---- accumulators (LHMM.py:149-161) ----------------------------------------------------------
    def add_acc(self, ksai_value, gamma_value):
        """Element-wise log-add into the log-domain transition accumulators."""
        self.__ksai_acc = np.logaddexp(self.__ksai_acc, np.asarray(ksai_value, dtype=np.float64))
        self.__gamma_acc = np.logaddexp(self.__gamma_acc, np.asarray(gamma_value, dtype=np.float64).reshape(-1))

    # ---- K1: scoring (LHMM.py:163-187) -------------------------------------------------------
    def cal_observation_pro(self, data, data_t, normalize=False, standard=False):
        """B[s, t] = log p(x_t | state s) for every utterance in `data`; rows of virtual states are
        log(1) = 0 (entry) and log(0) = -inf (exit) (AcousticModel.py:218-222,1039-1043)."""
        if normalize or standard:
            raise UnsupportedModel("normalize / standard scoring is not used by AcousticModel")
        eng = get_engine()
        gmms = [g for g in self.__profunction if isinstance(g, Clustering.GMM)]
        mix = gmms[0].mixture
        mean = torch.as_tensor(np.concatenate([g.mean for g in gmms])).to(eng.device)
        var = torch.as_tensor(np.concatenate([g.variance for g in gmms])).to(eng.device)
        alpha = torch.as_tensor(np.concatenate([g.alpha for g in gmms])).to(eng.device)
        W = eng.pack_gmm(mean, var, alpha, mix=0)
        out = []
        for d in range(len(data)):
            X = np.asarray(data[d], dtype=np.float64)[:data_t[d]]
            if X.shape[1] != gmms[0].dimension:
                from .Clustering import DataDimensionError
                raise DataDimensionError("frames have %d dimensions, the model %d" % (X.shape[1], gmms[0].dimension))
            rows = eng.prepare_rows(torch.as_tensor(X).to(eng.device))
            sc = eng.score_dense(rows, W, len(gmms), mix).double().cpu().numpy()  # [T, n_gmm]
            B = np.empty((len(self.__profunction), len(X)))
            k = 0
            for i, f in enumerate(self.__profunction):
                if isinstance(f, Clustering.GMM):
                    B[i] = sc[:, k]
                    k += 1
                else:
                    B[i] = f.point(None, log=True)
            out.append(B)
        self.__result_p = out

    # ---- K2 + K3: Baum-Welch on the sentence HMM (LHMM.py:526-544, 473-507) ---------------------
    def baulm_welch(self, show_q=False):
        """Forward-backward with the reference's pi-iteration (threshold 0.64) and update_acc:
        the unit HMMs in hmm_list receive the transition accumulators (add_acc) and their GMMs the
        occupancy / first / second moment statistics."""
        from .engine import Corpus, EStep, Model

        if self.__result_p is None or self.__hmm_list is None or self.__hmm_list == [self]:
            raise UnsupportedModel("baulm_welch needs probmat=[B] and hmm_list (the sentence HMM of "
                                   "AcousticModel.embedded); the profunc-only form crashes in the reference (Q16)")
        if self.__statesnum != 5:
            raise UnsupportedModel("the kernels cover state_num = 5 (3 emitting states per unit)")
        if self.__fix_list[2]:
            raise UnsupportedModel("fix_code bit 1 (pi fixed): the kernel always runs the reference's pi re-estimation "
                                   "loop (LHMM.py:447-471); AcousticModel never fixes pi")
        if self.__data is not None and len(self.__data) > 1:
            raise UnsupportedModel("several data items in one sentence HMM: the reference re-estimates ONE pi from all of "
                                   "them jointly inside the iteration (LHMM.py:455-466); AcousticModel passes one utterance "
                                   "per sentence HMM (AcousticModel.py:907-912) - use one LHMM per utterance or the batched "
                                   "AcousticModel.embedded_training")
        eng = get_engine()
        hl = self.__hmm_list
        L = len(hl)
        N = EMIT * L + 2
        tm = _banded_units(self.__transmat)
        if len(tm) != L:
            raise UnsupportedModel("transmat / hmm_list size mismatch")
        if not np.allclose(self.__pi, 1.0 / N):
            raise UnsupportedModel("the sentence HMM starts from the uniform pi of AcousticModel.embedded (Q4)")
        total_logp, iters = 0.0, []
        for d in range(len(self.__data)):
            X = np.asarray(self.__data[d], dtype=np.float64)
            T = int(self.__t[d]) if d < len(self.__t) else len(X)
            X = X[:T]
            B = np.asarray(self.__result_p[d], dtype=np.float64)
            if B.shape != (N, T) or np.any(B[0] != 0.0) or np.any(B[-1] != -np.inf):
                raise UnsupportedModel("probmat must be [3L+2, T] with entry row log(1) and exit row log(0)")
            gm = [[g for g in h.profunction[1:-1]] for h in hl]
            mean = np.stack([np.stack([g.mean for g in row]) for row in gm])
            var = np.stack([np.stack([g.variance for g in row]) for row in gm])
            alpha = np.stack([np.stack([g.alpha for g in row]) for row in gm])
            corpus = Corpus(eng, [np.arange(L, dtype=np.int32)], np.array([T], dtype=np.int32), L)
            model = Model(eng, mean, var, alpha, tm)
            es = EStep(eng, corpus, model)
            es.load_frames(torch.as_tensor(X).to(eng.device))
            model.pack(es.shift, es.inv_scale)
            corpus.emission_view(es.b, 0).copy_(torch.as_tensor(B[1:-1], dtype=torch.float32))
            es.forward_backward()
            skip_gmm = self.__fix_list[1]
            if not skip_gmm:
                es.accumulate()
            torch.cuda.synchronize()
            logp = float(es.utt_logp[0])
            total_logp += logp
            iters.append(int(es.utt_iters[0]))
            pt = es.pair_trans.double().cpu().numpy().reshape(L, EMIT, 3)
            lg = corpus.emission_view(es.lgam, 0).double().cpu().numpy()  # [3L, T] log gamma
            if not skip_gmm:
                occ, sx, sxx = es.linear_stats()
            with np.errstate(over="ignore"):
                socc = np.exp(lg).sum(axis=1).reshape(L, EMIT)
            for p, h in enumerate(hl):
                if not self.__fix_list[0]:
                    ks = np.full((EMIT, 5), -np.inf)
                    for r in range(EMIT):
                        ks[r, r + 1] = logp + pt[p, r, 0]
                        ks[r, r + 2] = logp + pt[p, r, 1]
                    h.add_acc(ks, logp + pt[p, :, 2])
                if not skip_gmm:
                    for r, g in enumerate(h.profunction[1:-1]):
                        g._add_linear(occ[p, r], socc[p, r], sx[p, r], sxx[p, r])
            if not self.__fix_list[2]:
                # pi := gamma_0 (LHMM.py:447-452); the entry state holds what the emitting states do not
                pi = np.zeros(N)
                pi[1:-1] = np.exp(lg[:, 0])
                pi[0] = max(0.0, 1.0 - pi[1:-1].sum())
                self.__pi = pi
            if show_q:
                self.log.note("log P(O) = %f after %d iterations" % (logp, iters[-1]), cls="i")
        self.log_likelihood = total_logp
        self.iterations = iters

    def update_acc(self):
        """LHMM.py:473-507 runs inside baulm_welch in this implementation (one fused device pass)."""

    # ---- M-step (LHMM.py:509-524) ------------------------------------------------------------
    def update_param(self, show_q=False, show_a=False, c_covariance=1e-3):
        """transmat[1:-1] = exp(ksai_acc - gamma_acc[:,None]); every emitting GMM re-estimates."""
        eng = get_engine()
        dev = eng.device
        if not self.__fix_list[0]:
            tmax = np.full((1, _eng.SLOTS), -np.inf)
            for r in range(EMIT):
                tmax[0, 3 * r + 0] = self.__ksai_acc[r, r + 1]
                tmax[0, 3 * r + 1] = self.__ksai_acc[r, r + 2]
                tmax[0, 3 * r + 2] = self.__gamma_acc[r]
            tsum = np.where(np.isfinite(tmax), 1.0, 0.0)
            tm = torch.as_tensor(np.asarray(self.__transmat, dtype=np.float64)[None].copy()).to(dev)
            M, D = 1, 1
            dummy = torch.zeros((1, 3, 1, _eng.KA), dtype=torch.float64, device=dev)
            one = torch.ones((1, 3, 1, 1), dtype=torch.float64, device=dev)
            # every device buffer stays referenced until the result has been read back: a temporary
            # freed right after _p() would hand its block to the next allocation of the same size
            tmax_d, tsum_d = torch.as_tensor(tmax).to(dev), torch.as_tensor(tsum).to(dev)
            one_v, one_a = one.clone(), one.clone().view(1, 3, 1)
            _eng.nat.call("pc_update_params", eng.h, 1, M, D, _eng._p(dummy), _eng._p(tmax_d), _eng._p(tsum_d),
                          None, None, float(c_covariance), 2, _eng._p(one), _eng._p(one_v), _eng._p(one_a),
                          _eng._p(tm), _eng._stream())
            self.__transmat = tm[0].cpu().numpy()
            del tmax_d, tsum_d, one_v, one_a
            if show_a:
                self.log.note("transmat:\n%s" % self.__transmat, cls="i")
        if not self.__fix_list[1] and self.__profunction is not None:
            for g in self.__profunction[1:-1]:
                g.update_param(show_q=show_q, c_covariance=c_covariance)
    # ---- K4: Viterbi (LHMM.py:546-609) ---------------------------------------------------------
    @staticmethod
    def viterbi(log, states, transmat, prob, pi, convert=False, end_state_back=False, show_mark_state=False):
        """fp64 max-plus recurrence on the device; ties to the lower state index, end state = first
        argmax over all states.  Returns (score, float64 array of state indices) or, with
        `convert`, the per-frame unit labels (AcousticModel.py:1026).  `end_state_back=True` has a
        stale-index bug in the reference (Q12) and is not reproduced."""
        from .engine import Corpus, host_log_bands, viterbi as _viterbi

        if end_state_back:
            raise UnsupportedModel("end_state_back=True (Q12) is not used by AcousticModel")
        eng = get_engine()
        B = np.asarray(prob, dtype=np.float64)
        N, T = B.shape
        tm = _banded_units(transmat)
        L = len(tm)
        if np.any(B[0] != 0.0) or np.any(B[-1] != -np.inf):
            raise UnsupportedModel("emission rows of the entry / exit states must be log(1) / log(0)")
        corpus = Corpus(eng, [np.arange(L, dtype=np.int32)], np.array([T], dtype=np.int32), L)
        ls, ln = host_log_bands(tm, eng.device)
        with np.errstate(divide="ignore"):
            logpi = torch.as_tensor(np.log(np.asarray(pi, dtype=np.float64))).to(eng.device)
        b64 = torch.zeros((corpus.emis_floats,), dtype=torch.float64, device=eng.device)
        corpus.emission_view(b64, 0).copy_(torch.as_tensor(B[1:-1]))
        score, path, _ = _viterbi(eng, corpus, b64, ls, ln, state_logpi=logpi, want_units=False)
        sc = float(score[0])
        mark = path.cpu().numpy().astype(np.float64)
        if show_mark_state and log is not None:
            log.note("viterbi states: %s" % mark, cls="i")
        if convert:
            return sc, np.array([states[int(s)] for s in mark])
        return sc, mark
