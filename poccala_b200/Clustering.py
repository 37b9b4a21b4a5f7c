"""Device-backed mirror of the reference's StatisticalModel/Clustering.py surface for the E-step
hot path: `Clustering.GMM` (scoring, Baum-Welch accumulators, re-estimation) and
`Clustering.ClusterInitialization` (k-means initialisation).  Same class / method / property
names and argument meaning as the reference (Clustering.py:36-767, 773-1044); every number is
produced by the sm_100a kernels behind include/poccala_b200.h - there is no CPU path.

Differences a maintainer should know:
  * accumulators are kept in the linear domain (sum gamma, sum gamma x, sum gamma (x-mu)^2) and the
    log-domain views the reference exposes (`acc`, `alpha_acc`, `mean_acc`) are derived on access;
  * `covariance` is the reference's [M,D,D] stack of diagonal matrices; only the diagonal is used
    (the reference extracts `.diagonal()` in util.gaussian_function, util.py:20-36);
  * the stand-alone `em()` trainer (Clustering.py:583-651,695-719) runs its E-step on the device; the SMEM
    split / merge search (Clustering.py:371-577) is host-side model selection over device-computed posteriors.
"""
from __future__ import annotations

import random as _random

import numpy as np
import torch

from . import engine as _eng
from .runtime import NullLog, get_engine


class DataDimensionError(Exception):
    """Exceptions.DataDimensionError (Exceptions.py:76): raised by GMM.point on a wrong-length frame."""


class DataUnLoadError(Exception):
    """Exceptions.DataUnLoadError: em() without data (Clustering.py:698-699)."""


class ParameterFileExistsError(Exception):
    """Exceptions.ParameterFileExistsError: the GMM_<id> parameter directory is missing (Clustering.py:295)."""


def _diag_stack(var):
    var = np.asarray(var, dtype=np.float64)
    M, D = var.shape
    out = np.zeros((M, D, D))
    idx = np.arange(D)
    out[:, idx, idx] = var
    return out


class Clustering(object):
    class LogInfoPrint(NullLog):
        """Clustering.LogInfoPrint stand-in (console logger of the reference)."""

    class GMM(object):
        def __init__(self, log=None, dimension=1, mix_level=1, data=None, alpha=None, mean=None, variance=None,
                     covariance=None, differentiation=True, gmm_id=0):
            """Clustering.py:37-104 (same arguments and defaults)."""
            self.log = log if log else Clustering.LogInfoPrint()
            self.__dimension = dimension
            self.__mix_level = mix_level
            self.__data = np.array(data) if data else None
            if mean is None:
                mean = np.random.random((mix_level, dimension)) if differentiation else np.zeros((mix_level, dimension))
            self.__mean = np.asarray(mean, dtype=np.float64)
            if covariance is not None:
                self.__covariance = np.asarray(covariance, dtype=np.float64)
            elif variance is not None:
                self.__covariance = _diag_stack(variance)
            elif differentiation:
                self.__covariance = _diag_stack(np.repeat(np.random.random((1, dimension)), mix_level, axis=0))
            else:
                self.__covariance = _diag_stack(np.ones((mix_level, dimension)))
            self.__alpha = (np.ones((mix_level,)) / mix_level) if alpha is None else np.asarray(alpha, dtype=np.float64)
            self.__bias = 100.
            self.__gmm_id = gmm_id
            self.__record = []
            self._clear_acc()

        # ---- parameter files (Clustering.py:234-255, 288-312): same directory layout and array
        # shapes as the reference, so models trained by either side load in the other.  (The
        # accumulator files of save_acc / init_acc are replaced by device reductions.)
        def save_parameter(self, path):
            import configparser
            import os

            path = path + '/GMM_%d' % self.__gmm_id
            if not os.path.exists(path):
                os.mkdir(path)
            np.save(path + '/GMM_means.npy', self.__mean)              # [M, D]
            np.save(path + '/GMM_covariance.npy', self.__covariance)   # [M, D, D]
            np.save(path + '/GMM_weight.npy', self.__alpha)            # [M]
            with open(path + '/GMM_config.ini', 'w+') as f:
                cfg = configparser.ConfigParser()
                cfg.add_section('Configuration')
                cfg.set('Configuration', 'MIXTURE', value=str(self.__mix_level))
                cfg.set('Configuration', 'DIMENSION', value=str(self.__dimension))
                cfg.set('Configuration', 'BIAS', value=str(self.__bias))
                cfg.write(f)

        def init_parameter(self, path):
            import os

            path = path + '/GMM_%d' % self.__gmm_id
            if not os.path.exists(path):
                raise ParameterFileExistsError(path)
            self.__mean = np.load(path + '/GMM_means.npy')
            self.__covariance = np.load(path + '/GMM_covariance.npy')
            self.__alpha = np.load(path + '/GMM_weight.npy')
            # the reference hands ConfigParser.read() an open file object, so the .ini never parses
            # and mixture / dimension / bias keep their constructor values (Q15); same here

        # ---- accumulator files (Clustering.py:257-283, 314-367): the reference's four log-domain arrays per
        # save - acc [M] = log sum gamma(j,m), alpha_acc = log sum gamma(j), mean_acc [M,D] = log sum gamma (x + 100),
        # covariance_acc [M,D] = log sum gamma (x - mu_old)^2 - derived from the linear statistics on save and
        # folded back into them on load (log-sum-exp over files == sum of the linear values).
        def save_acc(self, path):
            import os

            from .LHMM import LHMM

            path = path + '/GMM_%d' % self.__gmm_id
            dirs = [path, path + '/acc', path + '/alpha-acc', path + '/mean-acc', path + '/covariance-acc']
            for d in dirs:
                if not os.path.exists(d):
                    os.mkdir(d)
            np.save(LHMM._acc_file(dirs[1], 'GMM_acc'), self.acc)
            np.save(LHMM._acc_file(dirs[2], 'GMM_alpha_acc'), np.asarray(self.alpha_acc))
            np.save(LHMM._acc_file(dirs[3], 'GMM_mean_acc'), self.mean_acc)
            np.save(LHMM._acc_file(dirs[4], 'GMM_covariance_acc'), np.stack(self.covariance_acc_log))

        def init_acc(self, path):
            import os

            path = path + '/GMM_%d' % self.__gmm_id

            def files(sub):
                d = path + '/' + sub
                return [np.load(os.path.join(d, f)) for f in sorted(os.listdir(d))] if os.path.exists(d) else []

            with np.errstate(over="ignore", invalid="ignore"):
                occ_files = np.zeros_like(self._occ)
                for a in files('acc'):
                    occ_files = occ_files + np.exp(np.asarray(a, dtype=np.float64).reshape(-1))
                self._occ = self._occ + occ_files
                for a in files('alpha-acc'):
                    self._socc += float(np.exp(np.asarray(a, dtype=np.float64)).sum())
                # mean_acc carries the +100 bias of every contributing frame: sum gamma x = sum_f exp(mean_acc_f) - 100 sum_f occ_f
                m_files = files('mean-acc')
                if m_files:
                    self._sx = self._sx + sum(np.exp(np.asarray(m, dtype=np.float64)) for m in m_files) \
                        - self.__bias * occ_files[:, None]
                for c in files('covariance-acc'):
                    self._scc = self._scc + np.exp(np.asarray(c, dtype=np.float64))

        # ---- parameters (Clustering.py:122-229) -------------------------------------------
        @property
        def mean(self):
            return self.__mean

        @mean.setter
        def mean(self, value):
            self.__mean = np.asarray(value, dtype=np.float64)

        @property
        def covariance(self):
            return self.__covariance

        @covariance.setter
        def covariance(self, value):
            self.__covariance = np.asarray(value, dtype=np.float64)

        @property
        def variance(self):
            """[M,D] diagonal of `covariance` (what the scoring arithmetic uses)."""
            return np.ascontiguousarray(np.diagonal(self.__covariance, axis1=1, axis2=2))

        @property
        def alpha(self):
            return self.__alpha

        @alpha.setter
        def alpha(self, value):
            self.__alpha = np.asarray(value, dtype=np.float64)

        @property
        def dimension(self):
            return self.__dimension

        @property
        def mixture(self):
            return self.__mix_level

        @mixture.setter
        def mixture(self, mix_level):
            """Clustering.py:156-159: a new mixture count; the caller sets mean / covariance / alpha next."""
            self.__mix_level = int(mix_level)
            self._em_es = None
            self._clear_acc()

        @property
        def data(self):
            return self.__data

        @data.setter
        def data(self, data):
            """Clustering.py:166-174."""
            self.__data = data if isinstance(data, torch.Tensor) else np.array(data)
            self._em_es = None

        @property
        def bias(self):
            return self.__bias

        @property
        def gmm_id(self):
            return self.__gmm_id

        def add_data(self, data):
            """Clustering.py:231 (DataInitialization.add_data).  A [n, D] CUDA tensor (the rows
            engine.group_frames gathered) is taken as is and stays on the device."""
            if isinstance(data, torch.Tensor):
                self.__data = data if self.__data is None else torch.cat([torch.as_tensor(self.__data).to(data.device), data])
            else:
                d = np.array(data)
                self.__data = d if self.__data is None else np.append(self.__data, d, axis=0)
            self._em_es = None

        def clear_data(self):
            self.__data = None
            self._em_es = None

        # ---- accumulators: log-domain views of the linear statistics (Clustering.py:98-101) ----
        def _clear_acc(self):
            M, D = self.__mix_level, self.__dimension
            self._occ = np.zeros(M)
            self._socc = 0.0
            self._sx = np.zeros((M, D))
            self._scc = np.zeros((M, D))  # sum gamma (x - mu_old)^2 (Q8)

        def _add_linear(self, occ, socc, sx, sxx):
            """Fold statistics of one accumulation call in: occ [M], socc scalar, sx / sxx [M,D]
            (raw moments); the centred second moment uses the CURRENT mean (Clustering.py:677)."""
            mu = self.__mean
            self._occ += occ
            self._socc += float(socc)
            self._sx += sx
            self._scc += sxx - 2.0 * mu * sx + mu * mu * occ[:, None]

        @property
        def acc(self):
            with np.errstate(divide="ignore"):
                return np.log(self._occ)

        @property
        def alpha_acc(self):
            with np.errstate(divide="ignore"):
                return np.log(self._socc)

        @property
        def mean_acc(self):
            with np.errstate(divide="ignore", invalid="ignore"):
                return np.log(self._sx + self.__bias * self._occ[:, None])

        @property
        def covariance_acc(self):
            return self.__bias  # Q14: the reference's getter returns the bias (Clustering.py:206-209)

        @property
        def covariance_acc_log(self):
            """What the reference keeps in its private __covariance_acc: log sum gamma (x-mu_old)^2."""
            with np.errstate(divide="ignore"):
                return [np.log(np.maximum(r, 0.0)) for r in self._scc]

        # ---- scoring (Clustering.py:740-767 -> util.py:20-36) -------------------------------
        def _pack(self, eng):
            mean = torch.as_tensor(self.__mean).to(eng.device)
            var = torch.as_tensor(self.variance).to(eng.device)
            alpha = torch.as_tensor(self.__alpha).to(eng.device)
            return eng.pack_gmm(mean, var, alpha, mix=0)

        def score_frames(self, X):
            """log p(x_t) for every row of X [T,D] (batched `point(log=True)`): K1 dense kernel."""
            X = np.asarray(X, dtype=np.float64)
            if X.ndim != 2 or X.shape[1] != self.__dimension:
                raise DataDimensionError("expected frames of dimension %d" % self.__dimension)
            eng = get_engine()
            W = self._pack(eng)
            rows = eng.prepare_rows(torch.as_tensor(X).to(eng.device))
            out = eng.score_dense(rows, W, 1, self.__mix_level)
            return out[:, 0].double().cpu().numpy()

        def point(self, x, log=False, standard=False, record=False):
            """Score one frame.  `record=True` remembers the frame so that a following
            update_acc can rebuild the per-component posteriors (the reference stores the
            M component scores, Clustering.py:759-760)."""
            x = np.asarray(x, dtype=np.float64).reshape(-1)
            if len(x) != self.__dimension:
                raise DataDimensionError("frame has %d dimensions, the model %d" % (len(x), self.__dimension))
            if standard:
                raise NotImplementedError("standard=True (unit-Gaussian scoring) is outside the E-step path")
            lp = float(self.score_frames(x[None, :])[0])
            if record:
                self.__record.append(x)
            return lp if log else float(np.exp(lp))

        def gmm(self, x, mean, covariance, alpha, log=False, standard=False):
            """Clustering.py:721-738: mixture density with explicit parameters."""
            g = Clustering.GMM(self.log, self.__dimension, len(alpha), alpha=alpha, mean=mean, covariance=covariance)
            return g.point(x, log=log, standard=standard)

        # ---- Baum-Welch accumulation (Clustering.py:653-680) --------------------------------
        def update_acc(self, l_value, b_value, o_value):
            """l_value[T] = log gamma_t of this state, b_value[T] = its emission log-likelihoods,
            o_value[T,D] = the frames.  Runs the K3 contraction on the device for a one-state
            pseudo unit and folds the result into this GMM's accumulators; clears the record."""
            from .engine import Corpus, EStep, Model

            X = np.asarray(o_value, dtype=np.float64)
            T = len(X)
            l_value = np.asarray(l_value, dtype=np.float64).reshape(-1)
            b_value = np.asarray(b_value, dtype=np.float64).reshape(-1)
            assert len(l_value) == T and len(b_value) == T
            eng = get_engine()
            M, D = self.__mix_level, self.__dimension
            mean = np.broadcast_to(self.__mean, (1, 3, M, D)).copy()
            var = np.broadcast_to(self.variance, (1, 3, M, D)).copy()
            alpha = np.broadcast_to(self.__alpha, (1, 3, M)).copy()
            tm = np.zeros((1, 5, 5))
            corpus = Corpus(eng, [np.zeros(1, dtype=np.int32)], np.array([T], dtype=np.int32), 1)
            model = Model(eng, mean, var, alpha, tm)
            es = EStep(eng, corpus, model)
            es.load_frames(torch.as_tensor(X).to(eng.device))
            model.pack(es.shift, es.inv_scale)
            sp = _eng.nat.lib().pc_corpus_emission_floats(corpus.c) // T
            b = torch.zeros((T, sp), dtype=torch.float32, device=eng.device)  # unused states: b = 0, gamma = 0
            lg = torch.full((T, sp), float("-inf"), dtype=torch.float32, device=eng.device)
            b[:, 0] = torch.as_tensor(b_value, dtype=torch.float32)
            lg[:, 0] = torch.as_tensor(l_value, dtype=torch.float32)
            es.b.copy_(b.reshape(-1))
            es.lgam.copy_(lg.reshape(-1))
            es.accumulate()
            occ, sx, sxx = es.linear_stats()
            with np.errstate(over="ignore"):
                socc = float(np.exp(l_value[np.isfinite(l_value)]).sum())
            self._add_linear(occ[0, 0], socc, sx[0, 0], sxx[0, 0])
            self.__record = []

        # ---- M-step (Clustering.py:682-693) ---------------------------------------------------
        def update_param(self, show_q=False, c_covariance=1e-3):
            """alpha = occ/socc, mean = sx/occ, var = max(E[(x-mu_old)^2], c_covariance)."""
            eng = get_engine()
            M, D = self.__mix_level, self.__dimension
            dev = eng.device
            mu = self.__mean
            # pc_update_params consumes [G,80] rows (sum gamma x | sum gamma x^2) around the old mean
            sxx = self._scc + 2.0 * mu * self._sx - mu * mu * self._occ[:, None]
            acc = np.zeros((1, 3, M, _eng.KA))
            acc[0, :, :, :D] = self._sx
            acc[0, :, :, _eng.XS - 1] = self._occ
            acc[0, :, :, _eng.XS:_eng.XS + D] = sxx
            acc[0, :, :, _eng.KA - 1] = self._occ
            # alpha's denominator is the state occupancy (alpha_acc); the kernel derives it from the
            # component occupancies, which is the same number (sum_m gamma(j,m) = gamma(j))
            mean = torch.as_tensor(np.broadcast_to(mu, (1, 3, M, D)).copy()).to(dev)
            var = torch.as_tensor(np.broadcast_to(self.variance, (1, 3, M, D)).copy()).to(dev)
            alpha = torch.as_tensor(np.broadcast_to(self.__alpha, (1, 3, M)).copy()).to(dev)
            tm = torch.zeros((1, 5, 5), dtype=torch.float64, device=dev)
            tmax = torch.full((1, _eng.SLOTS), float("-inf"), dtype=torch.float64, device=dev)
            tsum = torch.zeros((1, _eng.SLOTS), dtype=torch.float64, device=dev)
            acc_d = torch.as_tensor(acc).to(dev)  # referenced until the results are read back
            _eng.nat.call("pc_update_params", eng.h, 1, M, D, _eng._p(acc_d), _eng._p(tmax),
                          _eng._p(tsum), None, None, float(c_covariance), 4, _eng._p(mean), _eng._p(var),
                          _eng._p(alpha), _eng._p(tm), _eng._stream())
            self.__mean = mean[0, 0].cpu().numpy()
            self.__covariance = _diag_stack(var[0, 0].cpu().numpy())
            self.__alpha = alpha[0, 0].cpu().numpy()
            if self._socc > 0.0:
                # the reference divides by alpha_acc, the state occupancy it was GIVEN (Clustering.py:685); the kernel
                # uses the sum of the component occupancies, which differs when b_value is not the exact
                # log-sum-exp of the component scores
                self.__alpha = self._occ / self._socc
            if show_q:
                self.log.note("GMM %s re-estimated" % self.__gmm_id, cls="i")
            # (the accumulators stay, as in the reference; AcousticModel builds fresh unit objects per M-step)

        # ---- stand-alone EM of one GMM (mode 1, Clustering.py:583-651, 695-719) ----------
        # The data set is laid out as a corpus of one pseudo unit whose first state carries this GMM and
        # whose posterior is 1 for every frame, so expectation() is K1 (scoring) + K3 (accumulation):
        # the component posteriors gamma_ik never leave the SM, what comes back are the sufficient
        # statistics occ_k = sum_i gamma_ik, sx_k = sum_i gamma_ik x_i, sxx_k = sum_i gamma_ik x_i^2.
        # maximization() and q_function() are closed forms of those (fp64, host: M x D numbers).
        def _em_setup(self):
            if getattr(self, "_em_es", None) is not None:
                return
            data = self.__data
            if data is not None and not isinstance(data, torch.Tensor):
                data = np.asarray(data, dtype=np.float64)
            if data is None or len(data) == 0:
                raise DataUnLoadError("GMM.em: no data loaded (Exceptions.DataUnLoadError)")
            if data.ndim != 2 or data.shape[1] != self.__dimension:
                raise DataDimensionError("expected [n, %d] data" % self.__dimension)
            eng = get_engine()
            n, chunk = len(data), 384
            n_frames = np.array([min(chunk, n - o) for o in range(0, n, chunk)], dtype=np.int32)
            corpus = _eng.Corpus(eng, [np.zeros(1, dtype=np.int32)] * len(n_frames), n_frames, 1)
            M, D = self.__mix_level, self.__dimension
            model = _eng.Model(eng, np.zeros((1, 3, M, D)), np.ones((1, 3, M, D)), np.ones((1, 3, M)) / M,
                               np.zeros((1, 5, 5)))
            es = _eng.EStep(eng, corpus, model)
            x_dev = torch.as_tensor(data).to(eng.device)
            es.load_frames(x_dev)
            # every frame belongs to state 0 with posterior 1; states 1, 2 of the pseudo unit are unused
            es.lgam.fill_(float("-inf"))
            es.lgam.view(-1, 8)[:, 0] = 0.0
            self._em_es, self._em_n = es, n
            self._em_x, self._em_rows, self._em_prev = x_dev, None, None

        def expectation(self):
            """Clustering.py:583-600: gamma_ik = alpha_k N(x_i; k) / sum_k' (...), reduced on the device
            to (occ, sx, sxx)."""
            self._em_setup()
            es, m = self._em_es, self._em_es.model
            dev = es.engine.device
            # the posteriors of this pass belong to THESE parameters (the split / merge search ranks on them)
            self._em_prev = (np.array(self.__mean), np.array(self.variance), np.array(self.__alpha))
            for r in range(3):
                m.mean[0, r].copy_(torch.as_tensor(self.__mean).to(dev))
                m.var[0, r].copy_(torch.as_tensor(self.variance).to(dev))
                m.alpha[0, r].copy_(torch.as_tensor(self.__alpha).to(dev))
            es.score()
            es.accumulate()
            occ, sx, sxx = es.linear_stats()
            self._em_stats = (occ[0, 0], sx[0, 0], sxx[0, 0])

        def maximization(self, c_covariance=1e-3):
            """Clustering.py:619-651: mean_k = sum gamma x / sum gamma (the reference goes through
            log(x + 100)), covariance around the NEW mean with the floor c_covariance, alpha_k = occ_k / n."""
            occ, sx, sxx = self._em_stats
            new_mean = sx / occ[:, None]
            var = (sxx - 2.0 * new_mean * sx + new_mean * new_mean * occ[:, None]) / occ[:, None]
            var = np.where(var < c_covariance, c_covariance, var)
            new_alpha = occ / self._em_n
            return new_mean, [np.diag(v) for v in var], new_alpha

        def q_function(self):
            """Clustering.py:602-613 with the posteriors of the last expectation() and the CURRENT
            parameters: sum_k occ_k log alpha_k + sum_ik gamma_ik log N(x_i; k), the second term written
            out on the sufficient statistics (log N is the reference's: -1/2 sum(var), Q1)."""
            occ, sx, sxx = self._em_stats
            mu, var = self.__mean, self.variance
            with np.errstate(divide="ignore"):
                value_1 = np.sum(occ * np.log(self.__alpha))
            const = -0.5 * self.__dimension * np.log(2 * np.pi) - 0.5 * var.sum(axis=1)
            quad = ((sxx - 2.0 * mu * sx + mu * mu * occ[:, None]) / var).sum(axis=1)
            return float(value_1 + np.sum(occ * const - 0.5 * quad))

        def em(self, show_q=False, smem=False, c_covariance=1e-3):
            """Clustering.py:695-719: iterate while Q grows by more than 1.28; with `smem` the split / merge
            search runs once the plain iteration has converged and the loop goes on if it found a better
            model."""
            self._em_setup()
            self.iterations = 0
            self.smem_trace = []
            q_value = -float("inf")
            while True:
                self.log.note("GMM Q: %f" % q_value, cls="i", show_console=show_q)
                self.expectation()
                self.iterations += 1
                mean, cov, alpha = self.maximization(c_covariance=c_covariance)
                self.__mean, self.__covariance, self.__alpha = mean, np.stack(cov), alpha
                _q = self.q_function()
                if _q - q_value > 1.28:
                    q_value = _q
                    continue
                if not smem:
                    break
                new_q_value = self._smem(q_value, c_covariance=c_covariance)
                if new_q_value is False:
                    break
                q_value = new_q_value
            self.q_value = q_value

        # ---- split and merge (SMEM, Clustering.py:371-577) ----------------------------
        # The search itself is model selection on [M, n] posteriors and runs on the host in fp64, as SURVEY
        # section 8 f2 places it; every Gaussian evaluation it needs (the posteriors it ranks on, the three
        # candidate components, the Q terms) is one launch of the dense scoring kernel with one component per
        # output column.  The reference's arithmetic is kept as it is written, including where it mixes
        # domains: `__gamma` holds LOG posteriors after expectation(), and __J_merge, __J_split and the
        # `gamma_sum` of __SMEM use those logs as if they were weights (:381, :395-409, :536-540); __reestimate
        # then returns posterior x (sum of three logs) (:481), which maximization() and q_function() read as
        # log weights.  The accept / reject decision and the ranking are reproduced on exactly those numbers.
        def _component_scores(self, mean, var, alpha):
            """[K, n] fp64: log alpha_k + log N(x_i; k) (util.py:20-31 with log=True, Q1) on the device."""
            es = self._em_es
            eng = es.engine
            if self._em_rows is None:
                self._em_rows = eng.prepare_rows(self._em_x, es.shift, es.inv_scale)
            K = len(alpha)
            W = eng.pack_gmm(torch.as_tensor(np.ascontiguousarray(mean, dtype=np.float64)).to(eng.device),
                             torch.as_tensor(np.ascontiguousarray(var, dtype=np.float64)).to(eng.device),
                             torch.as_tensor(np.ascontiguousarray(alpha, dtype=np.float64)).to(eng.device),
                             es.shift, es.inv_scale, mix=0)
            out = eng.score_dense(self._em_rows, W, K, 1)
            return out.double().t().contiguous().cpu().numpy()

        def _log_posteriors(self):
            """The `__gamma` table the reference holds when __SMEM starts: log posteriors of the last
            expectation(), i.e. under the parameters BEFORE the last maximization (Clustering.py:591-599)."""
            s = self._component_scores(*self._em_prev)
            m = s.max(axis=0)
            return s - (m + np.log(np.exp(s - m).sum(axis=0)))

        def _j_merge(self, gamma):
            """Clustering.py:372-386: cosine of every pair of `__gamma` rows, best first."""
            M = len(gamma)
            out = []
            with np.errstate(all="ignore"):
                norm = [np.linalg.norm(gamma[i]) for i in range(M)]
                for i in range(M):
                    for j in range(i + 1, M):
                        out.append([i, j, float(np.dot(gamma[i], gamma[j])) / (norm[i] * norm[j])])
            out.sort(key=lambda r: r[2], reverse=True)
            return out

        def _j_split(self, gamma, X, mean, var):
            """Clustering.py:388-430.  For component k and point x: the members of k (points whose largest
            `__gamma` entry is k's, first one on ties) sorted by distance to x; p = sum_r (r / n_k) gamma_k[r-th]
            / sum_i gamma_k[i]; J_split(k) = sum_x p log(p / N_std(x; k)) with the `standard=True` density of
            util.py:24-26 (deviation multiplied by sqrt(var), unit covariance)."""
            M, n = gamma.shape
            D = X.shape[1]
            owner = np.argmax(gamma, axis=0)
            out = []
            with np.errstate(all="ignore"):
                for k in range(M):
                    idx = np.flatnonzero(owner == k)
                    nk = len(idx)
                    pg = gamma[k]
                    p2 = np.sum(pg)
                    if nk == 0:
                        p = np.zeros(n) / p2
                    else:
                        Y, pgk = X[idx], pg[idx]
                        rank = np.arange(nk) / nk
                        p1 = np.empty(n)
                        step = max(1, (1 << 22) // (nk * D))
                        for o in range(0, n, step):
                            diff = X[o:o + step, None, :] - Y[None, :, :]
                            order = np.argsort(np.einsum("abk,abk->ab", diff, diff), axis=1, kind="stable")
                            p1[o:o + step] = np.cumsum(rank * pgk[order], axis=1)[:, -1]
                        p = p1 / p2
                    dev = (X - mean[k]) * np.sqrt(var[k])
                    dens = np.exp(-0.5 * np.sum(dev * dev, axis=1)) / (2 * np.pi) ** (D / 2)
                    out.append([k, float(np.cumsum(p * np.log(p / dens))[-1])])
            out.sort(key=lambda r: r[1], reverse=True)
            return out, owner

        def _split(self, x, X, owner):
            """Clustering.py:443-465: 2-means of the points component x owns, centres jittered by
            numpy's global generator, spherical covariance det^(1/D), half the weight each."""
            pts = X[owner == x]
            if len(pts) < self.__mix_level:
                return False
            mean_, _, _, _ = Clustering.ClusterInitialization(list(pts), 2, self.__dimension, self.log).kmeans(algorithm=1)
            D = self.__dimension
            m1 = mean_[0] + np.random.rand(D) * 1e-2
            m2 = mean_[1] + np.random.rand(D) * 1e-2
            v = np.full(D, np.linalg.det(self.__covariance[x]) ** (1 / D))
            return [m1, m2], [v, v.copy()], [self.__alpha[x] * 0.5, self.__alpha[x] * 0.5]

        def _smem(self, q_value, c_max=5, c_covariance=1e-3):
            """Clustering.py:483-577.  Returns the new Q value when a merge(i, j) + split(k) candidate beats
            `q_value`, False otherwise.  As in the reference only the first candidate whose split is possible is
            evaluated (both branches of :563-577 return).  Where the reference cannot continue - an accepted
            candidate leaves `__mix_level` at M - 3 (:556) and the next maximization() raises for M > 4; no
            splittable candidate returns None into a float subtraction - the accepted model is kept with all
            its M components, and "nothing to split" ends the search."""
            if self.__mix_level < 3:
                return False
            M, D, n = self.__mix_level, self.__dimension, self._em_n
            X = self._em_x.cpu().numpy()
            mean, var, alpha = np.array(self.__mean), np.array(self.variance), np.array(self.__alpha)
            gamma = self._log_posteriors()
            merge_list = self._j_merge(gamma)
            split_list, owner = self._j_split(gamma, X, mean, var)
            candidates = []
            for i, j, _ in merge_list:
                for k, _ in split_list:
                    if k != i and k != j and len(candidates) < c_max:
                        candidates.append([i, j, k])
                if len(candidates) == c_max:
                    break
            trace = {"merge": merge_list, "split": split_list, "candidates": candidates}
            self.smem_trace.append(trace)
            for cand in candidates:
                i, j, k = cand
                sp = self._split(k, X, owner)
                if sp is False:
                    continue
                a_m = alpha[i] + alpha[j]
                mean3 = np.stack([(mean[i] * alpha[i] + mean[j] * alpha[j]) / a_m] + sp[0])
                var3 = np.stack([(var[i] * alpha[i] + var[j] * alpha[j]) / a_m] + sp[1])
                alpha3 = np.array([a_m] + sp[2])
                with np.errstate(all="ignore"):
                    gamma_sum = gamma[i] + gamma[j] + gamma[k]
                    s3 = self._component_scores(mean3, var3, alpha3)
                    m3 = s3.max(axis=0)
                    g3 = np.exp(s3 - (m3 + np.log(np.exp(s3 - m3).sum(axis=0)))) * gamma_sum  # __reestimate :467-481
                    # maximization() on g3 read as log weights (:619-651), in linear arithmetic
                    top = g3.max(axis=1, keepdims=True)
                    w = np.exp(g3 - top)
                    ws = w.sum(axis=1)
                    new_mean = (w @ X) / ws[:, None]
                    new_var = np.stack([(w[c] @ (X - new_mean[c]) ** 2) / ws[c] for c in range(3)])
                    new_var = np.where(new_var < c_covariance, c_covariance, new_var)
                    new_alpha = np.exp(top[:, 0]) * ws / n
                    # q_function() of the three new components (:602-613)
                    w3 = np.exp(g3)
                    ln3 = self._component_scores(new_mean, new_var, np.ones(3))
                    q_1 = float(np.sum(w3.sum(axis=1) * np.log(new_alpha)) + np.sum(w3 * ln3))
                    # ... and of the components that stay, with the table of the last expectation()
                    keep = np.array([c for c in range(M) if c not in cand], dtype=np.int64)
                    q_2 = 0.0
                    if len(keep):
                        wk = np.exp(gamma[keep])
                        lnk = self._component_scores(mean[keep], var[keep], np.ones(len(keep)))
                        q_2 = float(np.sum(wk.sum(axis=1) * np.log(alpha[keep])) + np.sum(wk * lnk))
                new_q = q_1 + q_2
                trace.update(chosen=cand, mean3=mean3, var3=var3, alpha3=alpha3, new_mean=new_mean, new_var=new_var,
                             new_alpha=new_alpha, q_1=q_1, q_2=q_2, accepted=bool(new_q > q_value))
                if new_q > q_value:
                    self.__mean = np.append(mean[keep], new_mean, axis=0)
                    self.__covariance = _diag_stack(np.append(var[keep], new_var, axis=0))
                    self.__alpha = np.append(alpha[keep], new_alpha)
                    return new_q
                return False
            return False

        def theta(self):
            """Clustering.py:721-726: {'theta_k': [mean_k, diag(covariance_k)]}."""
            return {"theta_%d" % i: [self.__mean[i], self.__covariance[i].diagonal()] for i in range(self.__mix_level)}

    class ClusterInitialization(object):
        def __init__(self, data, k, dimension, log=None):
            """Clustering.py:774-790: data = list (or array) of D-vectors, k clusters."""
            self.__data = data
            self.__k = k
            self.__dimension = dimension
            self.log = log if log else Clustering.LogInfoPrint()
            self.passes = None
            self.moves = None

        @staticmethod
        def cal_distance(d1, d2, arg=2):
            """Clustering.py:796-801 (Q2): the loop returns in its first iteration, so only
            dimension 0 contributes.  Scalar helper kept for API parity; the device kernel
            evaluates the same expression for every point."""
            return float((abs(d1[0] - d2[0]) ** arg) ** (1 / arg))

        @staticmethod
        def cal_variance(cluster, algorithm=None):
            """Clustering.py:807-832: per-dimension standard deviation of `cluster` = [centre,
            points] with the 1e-4 variance floor, computed by pc_kmeans_finish."""
            centre, points = cluster[0], cluster[1]
            if algorithm == "kmeans":
                points = list(points.values())
            pts = np.ascontiguousarray(points, dtype=np.float64)
            centre = np.asarray(centre, dtype=np.float64)
            eng = get_engine()
            n, D = pts.shape
            x = torch.as_tensor(pts).to(eng.device)
            off = np.array([0, n], dtype=np.int64)
            # pc_kmeans_run writes the problem table pc_kmeans_finish reads; one pass is enough here
            out = _eng.kmeans_run(eng, x, off, 1, np.zeros((1, 1), np.int32), max_passes=1)
            ml = torch.arange(n, dtype=torch.int32, device=eng.device)  # the caller's point order
            mc = torch.tensor([[n]], dtype=torch.int32, device=eng.device)
            mean = torch.empty((1, 1, D), dtype=torch.float64, device=eng.device)
            var = torch.empty_like(mean)
            al = torch.empty((1, 1), dtype=torch.float64, device=eng.device)
            _eng.nat.call("pc_kmeans_finish", eng.h, 1, _eng._p(off), _eng._p(x), D, 1, _eng._p(out["workspace"]),
                          _eng._p(ml), _eng._p(mc), _eng._p(mean), _eng._p(var), _eng._p(al), _eng._stream())
            if not np.allclose(mean[0, 0].cpu().numpy(), centre, rtol=1e-12, atol=1e-12):
                raise NotImplementedError("cal_variance about a centre other than the cluster mean is not used by "
                                          "the reference's k-means (Clustering.py:938-944)")
            return list(np.sqrt(var[0, 0].cpu().numpy()))

        def kmeans(self, algorithm=0, cov_matrix=False):
            """Clustering.py:838-1044 with algorithm=1 (k-means++ style seeding drawn from Python's
            `random`, greedy passes on the device).  Returns (mean[K,D], std[K,D] or cov[K,D,D],
            alpha list, clustered_data list of point lists in insertion order)."""
            if algorithm != 1:
                raise NotImplementedError("only algorithm=1 is used by AcousticModel (AcousticModel.py:499,553)")
            data = np.asarray(self.__data, dtype=np.float64)
            if data.ndim != 2 or data.shape[1] != self.__dimension:
                raise DataDimensionError("expected [n, %d] data" % self.__dimension)
            n, K = len(data), self.__k
            eng = get_engine()
            seeds = _eng.kmeans_seed_points(np.ascontiguousarray(data[:, 0]), K, _random)
            x = torch.as_tensor(data).to(eng.device)
            out = _eng.kmeans_run(eng, x, np.array([0, n]), K, np.array([seeds], dtype=np.int32))
            counts = out["member_count"][0].cpu().numpy()
            ml = out["member_list"].cpu().numpy()
            self.passes, self.moves = int(out["passes"][0]), int(out["moves"][0])
            for _ in range(self.moves):  # the reference draws one dict key per move (Clustering.py:932)
                _random.random()
            mean = out["mean"][0].cpu().numpy()
            var = out["var"][0].cpu().numpy()
            alpha = [float(a) for a in out["alpha"][0].cpu().numpy()]
            clustered, o = [], 0
            for kk in range(K):
                clustered.append([data[i] for i in ml[o:o + counts[kk]]])
                o += counts[kk]
            if cov_matrix:
                return mean, _diag_stack(var), alpha, clustered
            return mean, np.sqrt(var), alpha, clustered
