"""Host side of the E-step engine: device buffers (torch tensors), streams, the EM iteration and the
NCCL reduction of the accumulators.  Every number is produced by the CUDA kernels behind the C ABI
(poccala_b200/_native.py); torch is plumbing (memory, streams, torch.distributed).

Mirrors the work AcousticModel.embedded_training orchestrates per utterance / per unit
(AcousticModel.py:842-935): multi_embedded_training_1 (score -> sentence HMM -> Baum-Welch ->
accumulate) becomes three batched kernels over a resident corpus; the file-based accumulator
merge (LHMM.py:256-290, Clustering.py:314-367) becomes device reductions plus one NCCL allreduce;
multi_embedded_training_2 (update_param per unit) becomes one M-step kernel.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as nat
from . import distributed as _dist
from ._native import _p

EMIT, STATES, XS, KA, SLOTS = nat.EMIT, nat.STATES, nat.XS, nat.KA, nat.TRANS_SLOTS


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Engine:
    """One per GPU (pc_handle)."""

    def __init__(self, device=0):
        if not torch.cuda.is_available():
            raise RuntimeError("poccala_b200.Engine needs a CUDA device (B200); there is no CPU path")
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        h = C.c_void_p()
        nat.call("pc_create", int(device), C.byref(h))
        self.h = h

    def close(self):
        if self.h is not None:
            nat.lib().pc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key, value):
        nat.call("pc_set_option", self.h, key.encode(), int(value))

    def get_option(self, key):
        return int(nat.lib().pc_get_option(self.h, key.encode()))

    def side_stream(self):
        """Stream for work that may run beside the main kernels (transition reductions under K3)."""
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    @property
    def launches(self):
        return self.get_option("launches")

    # ------------------------------------------------------------------ buffers
    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def corpus(self, labels, n_frames, n_units):
        return Corpus(self, labels, n_frames, n_units)

    # ------------------------------------------------------------------ kernels
    def pack_gmm(self, mean, var, alpha, shift=None, inv_scale=None, out=None, mix=0):
        """mean/var [G,D] fp64 cuda, alpha [G] -> W buffer (pc_gmm_bytes(G) bytes: fp32 rows, scale,
        flags, per-unit fp16 tensor-core images; pc_pack_gmm).  mix = components per state of an
        acoustic model laid out [unit][3][mix]; 0 for a flat list of Gaussians."""
        G, D = mean.shape
        W = out if out is not None else self.empty((int(nat.lib().pc_gmm_bytes(G)),), torch.uint8)
        nat.call("pc_pack_gmm", self.h, _p(mean), _p(var), _p(alpha), _p(shift), _p(inv_scale), G, D, int(mix),
                 _p(W), _stream())
        return W

    def prepare_rows(self, x, shift=None, inv_scale=None):
        """x [F,D] fp32/fp64 cuda -> fp32 rows [F,40] (dense scoring sweep; pc_prepare_rows_*)."""
        F, D = x.shape
        if D > nat.DIM_MAX:
            raise ValueError("feature dimension %d exceeds %d" % (D, nat.DIM_MAX))
        if x.dtype not in (torch.float32, torch.float64):
            raise TypeError("frames must be float32 or float64")
        x = x.contiguous()
        X = self.empty((F, XS), torch.float32)
        fn = "pc_prepare_rows_f64" if x.dtype == torch.float64 else "pc_prepare_rows_f32"
        nat.call(fn, self.h, _p(x), F, D, _p(shift), _p(inv_scale), _p(X), _stream())
        return X

    def prepare_frames(self, corpus, x, shift=None, inv_scale=None, out=None):
        """x [F,D] fp32/fp64 cuda (utterances concatenated in corpus order) -> the corpus' X buffer
        (standardised fp32 rows + one fp16 tensor-core image per 128-frame tile; pc_prepare_frames_*)."""
        F, D = x.shape
        if D > nat.DIM_MAX:
            raise ValueError("feature dimension %d exceeds %d" % (D, nat.DIM_MAX))
        if F != corpus.total_frames:
            raise ValueError("expected %d frames, got %d" % (corpus.total_frames, F))
        if x.dtype not in (torch.float32, torch.float64):
            raise TypeError("frames must be float32 or float64")
        x = x.contiguous()
        X = out if out is not None else self.empty((int(nat.lib().pc_corpus_frames_bytes(corpus.c)),), torch.uint8)
        fn = "pc_prepare_frames_f64" if x.dtype == torch.float64 else "pc_prepare_frames_f32"
        nat.call(fn, self.h, corpus.c, _p(x), D, _p(shift), _p(inv_scale), _p(X), _stream())
        return X

    def score_dense(self, X, W, n_states, mix, out=None):
        F = X.shape[0]
        out = out if out is not None else self.empty((F, n_states), torch.float32)
        nat.call("pc_gmm_score_dense", self.h, _p(X), F, _p(W), n_states, mix, _p(out), _stream())
        return out


    def score_dense_tc(self, x, mean, var, alpha, mix, frames_per_group=384, shift=None, inv_scale=None):
        """Every frame against every state of ONE model (the GMM scoring sweep of BASELINE.json
        configs[2]; LHMM.cal_observation_pro's arithmetic, LHMM.py:163-187) on the tensor-core kernel:
        the frames are cut into groups of `frames_per_group`, each a pseudo utterance labelled with all
        ceil(S/3) pseudo units, so K1 streams the whole model past every group of three frame tiles.
        x [F,D] device tensor; mean / var [S*mix, D], alpha [S*mix] fp64 device tensors (S states).
        Returns b [F, S] fp32 (a view of the [F, SP] emission buffer)."""
        F, D = x.shape
        G = mean.shape[0]
        if G % mix:
            raise ValueError("number of Gaussians is not a multiple of mix")
        S = G // mix
        U = (S + EMIT - 1) // EMIT
        pad = U * EMIT * mix - G
        if pad:  # dead Gaussians (alpha = 0) fill the last pseudo unit
            mean = torch.cat([mean, torch.zeros((pad, D), dtype=mean.dtype, device=mean.device)])
            var = torch.cat([var, torch.ones((pad, D), dtype=var.dtype, device=var.device)])
            alpha = torch.cat([alpha, torch.zeros((pad,), dtype=alpha.dtype, device=alpha.device)])
        n_frames = np.full((F + frames_per_group - 1) // frames_per_group, frames_per_group, dtype=np.int32)
        if F % frames_per_group:
            n_frames[-1] = F % frames_per_group
        labels = np.ascontiguousarray(np.broadcast_to(np.arange(U, dtype=np.int32), (len(n_frames), U)))
        corpus = Corpus(self, labels, n_frames, U)
        W = self.pack_gmm(mean.contiguous(), var.contiguous(), alpha.contiguous(), shift, inv_scale, mix=mix)
        X = self.prepare_frames(corpus, x, shift, inv_scale)
        b = self.empty((corpus.emis_floats,), torch.float32)
        nat.call("pc_gmm_score", self.h, corpus.c, _p(X), _p(W), int(mix), _p(b), _stream())
        sp = (EMIT * U + 7) & ~7
        self._dense_keep = (corpus, X, W)  # alive until the stream has run
        return b.view(F, sp)[:, :S]


class Corpus:
    """Descriptor tables of a set of utterances (pc_corpus) + the resident frame matrix."""

    def __init__(self, engine, labels, n_frames, n_units):
        self.engine = engine
        n_frames = np.ascontiguousarray(n_frames, dtype=np.int32)
        if isinstance(labels, np.ndarray) and labels.ndim == 2:
            n_labels = np.full(labels.shape[0], labels.shape[1], dtype=np.int32)
            flat = np.ascontiguousarray(labels.reshape(-1), dtype=np.int32)
        else:
            n_labels = np.array([len(l) for l in labels], dtype=np.int32)
            flat = (np.ascontiguousarray(np.concatenate([np.asarray(l, dtype=np.int32) for l in labels]))
                    if len(labels) else np.zeros(0, np.int32))
        if len(n_labels) != len(n_frames):
            raise ValueError("labels / n_frames length mismatch")
        self.n_utt = len(n_frames)
        self.n_units = int(n_units)
        self.n_frames = n_frames
        self.n_labels = n_labels
        self.labels_flat = flat
        c = C.c_void_p()
        nat.call("pc_corpus_create", engine.h, self.n_utt, _p(n_frames), _p(n_labels), _p(flat), self.n_units,
                 C.byref(c))
        self.c = c
        l = nat.lib()
        self.total_frames = int(l.pc_corpus_total_frames(c))
        self.emis_floats = int(l.pc_corpus_emission_floats(c))
        self.n_pairs = int(l.pc_corpus_total_pairs(c))
        self.total_states = int(l.pc_corpus_total_states(c))
        offs = [np.empty(self.n_utt + 1, dtype=np.int64) for _ in range(4)]
        nat.call("pc_corpus_offsets", c, *[_p(o) for o in offs])
        self.frame_off, self.emis_off, self.pair_off, self.state_off = offs
        self.X = None

    def close(self):
        if self.c is not None:
            nat.lib().pc_corpus_destroy(self.c)
            self.c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def emission_view(self, buf, u):
        """[3L, T] view of utterance u inside a b / lgam buffer (stored time-major [T][SP])."""
        T, L = int(self.n_frames[u]), int(self.n_labels[u])
        sp = (EMIT * L + 7) & ~7
        return buf[self.emis_off[u]:self.emis_off[u + 1]].view(T, sp)[:, :EMIT * L].t()


class Model:
    """GMM-HMM parameters resident on the device (fp64, the reference's precision).
    mean/var [U,3,M,D], alpha [U,3,M], transmat [U,5,5]."""

    def __init__(self, engine, mean, var, alpha, transmat):
        dev = engine.device
        self.engine = engine
        self.mean = torch.as_tensor(np.asarray(mean), dtype=torch.float64).to(dev).contiguous()
        self.var = torch.as_tensor(np.asarray(var), dtype=torch.float64).to(dev).contiguous()
        self.alpha = torch.as_tensor(np.asarray(alpha), dtype=torch.float64).to(dev).contiguous()
        self.transmat = torch.as_tensor(np.asarray(transmat), dtype=torch.float64).to(dev).contiguous()
        self.n_units, _, self.mix, self.dim = self.mean.shape
        self.n_gauss = self.n_units * EMIT * self.mix
        self.W = engine.empty((int(nat.lib().pc_gmm_bytes(self.n_gauss)),), torch.uint8)

    def pack(self, shift=None, inv_scale=None):
        self.engine.pack_gmm(self.mean.view(self.n_gauss, self.dim), self.var.view(self.n_gauss, self.dim),
                             self.alpha.view(self.n_gauss), shift, inv_scale, out=self.W, mix=self.mix)
        return self.W

    def log_bands(self):
        """log of the diagonal / super-diagonal of every unit transmat, [U,5] each (device, fp64)."""
        if not hasattr(self, "_ls"):
            self._ls = self.engine.empty((self.n_units, STATES), torch.float64)
            self._ln = self.engine.empty((self.n_units, STATES), torch.float64)
        nat.call("pc_log_bands", self.engine.h, _p(self.transmat), self.n_units, _p(self._ls), _p(self._ln), _stream())
        return self._ls, self._ln

    def numpy(self):
        return (self.mean.cpu().numpy(), self.var.cpu().numpy(), self.alpha.cpu().numpy(),
                self.transmat.cpu().numpy())


class EStep:
    """Buffers and kernels of one EM iteration over a resident corpus."""

    def __init__(self, engine, corpus, model, standardise=True):
        self.engine, self.corpus, self.model = engine, corpus, model
        e = engine
        self.b = e.empty((max(corpus.emis_floats, 1),), torch.float32)
        self.lgam = e.empty((max(corpus.emis_floats, 1),), torch.float32)
        self.utt_logp = e.empty((corpus.n_utt,), torch.float64)
        self.utt_iters = e.empty((corpus.n_utt,), torch.int32)
        self.pair_trans = e.empty((max(corpus.n_pairs, 1), SLOTS), torch.float32)
        # linear GMM statistics and transition sums share one flat buffer: one allreduce(SUM)
        n_acc = model.n_gauss * KA
        self.flat = e.empty((n_acc + model.n_units * SLOTS,), torch.float64)
        self.acc = self.flat[:n_acc].view(model.n_gauss, KA)
        self.tsum = self.flat[n_acc:].view(model.n_units, SLOTS)
        self.tmax = e.empty((model.n_units, SLOTS), torch.float64)
        self._own = (self.flat, self.acc, self.tsum, self.tmax)
        self.shift = None
        self.inv_scale = None
        self.standardise = standardise

    # peer-memory reduction ----------------------------------------------------------------
    def use_peer(self, peer):
        """Keep the statistics in `peer`'s exchange block (distributed.PeerExchange): estep() then runs no
        collective at all, mstep() reduces over the ranks while it reads (pc_update_params_peer) and leaves
        the summed statistics in acc / tsum / tmax."""
        self.peer = peer
        peer._bound.add(self)
        self._peer_bind()

    def use_nccl(self):
        """Back to the statistics buffers of this object and the NCCL reduction (undoes use_peer)."""
        self.peer = None
        self.flat, self.acc, self.tsum, self.tmax = self._own
        self._acc_clean = False

    def _peer_bind(self):
        self.acc, self.tsum, self.tmax = self.peer.current()
        self.flat = None
        self._acc_clean = False

    # frames -------------------------------------------------------------------------------
    def load_frames(self, x, group=None, shift=None, inv_scale=None):
        """x: [F,D] float tensor on the device (utterances concatenated in corpus order).
        Standardisation uses the per-dimension mean / std of the WHOLE corpus: with a process group
        the moments are allreduced first, so every rank maps its frames (and therefore its
        accumulators, which live in the standardised space) the same way.  `shift` / `inv_scale`
        override the statistics (fp64 device tensors [D])."""
        if x.shape[0] != self.corpus.total_frames:
            raise ValueError("expected %d frames, got %d" % (self.corpus.total_frames, x.shape[0]))
        if shift is not None:
            self.shift, self.inv_scale = shift.contiguous(), inv_scale.contiguous()
        elif self.standardise:
            xd = x.to(torch.float64)
            mom = torch.cat([xd.sum(dim=0), (xd * xd).sum(dim=0),
                             torch.full((1,), float(x.shape[0]), dtype=torch.float64, device=x.device)])
            if group is not None:
                torch.distributed.all_reduce(mom, group=group)
            D = x.shape[1]
            n = mom[-1]
            mu = mom[:D] / n
            sd = (mom[D:2 * D] / n - mu * mu).clamp_min(0.0).sqrt().clamp_min(1e-12)
            self.shift = mu.contiguous()
            self.inv_scale = (1.0 / sd).contiguous()
        self.corpus.X = self.engine.prepare_frames(self.corpus, x, self.shift, self.inv_scale, out=self.corpus.X)
        clamped = self.engine.get_option("clamped")  # synchronises: once per corpus
        if clamped:
            raise ValueError("%d feature values lie more than 240 standard units from the shift and were clamped: "
                             "standardise the frames (EStep(standardise=True)) or pass matching shift / inv_scale"
                             % clamped)
        return self.corpus.X

    # kernels ------------------------------------------------------------------------------
    def score(self):
        m = self.model
        m.pack(self.shift, self.inv_scale)
        nat.call("pc_gmm_score", self.engine.h, self.corpus.c, _p(self.corpus.X), _p(m.W), m.mix, _p(self.b),
                 _stream())

    def log_bands_async(self):
        """log(transmat) bands and the clearing of the statistics buffer on the side stream: they depend
        on the model / on nothing, so they run beside K1 (forward_backward() joins)."""
        cur = torch.cuda.current_stream()
        side = self.engine.side_stream()
        fork = torch.cuda.Event()
        fork.record(cur)
        side.wait_event(fork)
        with torch.cuda.stream(side):
            self.acc.zero_()  # K3 accumulates into it; cleared here, beside K1, instead of in front of K3
            self._acc_clean = True
            self._bands = self.model.log_bands()
            self._bands_done = torch.cuda.Event()
            self._bands_done.record(side)

    def forward_backward(self):
        if getattr(self, "_bands_done", None) is not None:
            torch.cuda.current_stream().wait_event(self._bands_done)
            (ls, ln), self._bands_done = self._bands, None
        else:
            ls, ln = self.model.log_bands()
        nat.call("pc_forward_backward", self.engine.h, self.corpus.c, _p(self.b), _p(ls), _p(ln), _p(self.lgam),
                 _p(self.utt_logp), _p(self.utt_iters), _p(self.pair_trans), _stream())

    def accumulate(self):
        m = self.model
        if not getattr(self, "_acc_clean", False):
            self.acc.zero_()
        elif getattr(self, "_bands_done", None) is not None:  # cleared on the side stream, not joined yet
            torch.cuda.current_stream().wait_event(self._bands_done)
        self._acc_clean = False
        nat.call("pc_accumulate", self.engine.h, self.corpus.c, _p(self.corpus.X), _p(m.W), m.mix, _p(self.b),
                 _p(self.lgam), _p(self.acc), _stream())

    def reduce_transitions_async(self, group=None):
        """Per-unit log-sum-exp of the transition counts (pc_transitions_max / _sum) and, with a process
        group, the MAX collective between the two - on the engine's side stream, forked from the current
        stream: they depend on K2's outputs only, so they run under the accumulation kernel (small
        blocks fit beside its one persistent CTA per SM).  reduce_statistics() joins."""
        cur = torch.cuda.current_stream()
        side = self.engine.side_stream()
        fork = torch.cuda.Event()
        fork.record(cur)
        side.wait_event(fork)
        with torch.cuda.stream(side):
            nat.call("pc_transitions_max", self.engine.h, self.corpus.c, _p(self.utt_logp), _p(self.pair_trans),
                     _p(self.tmax), _stream())  # writes every (unit, slot): -inf where the unit has no pair
            if getattr(self, "peer", None) is None:  # (the peer M-step combines the ranks' (max, sum) pairs itself)
                _dist.allreduce_transition_maxima(self.tmax, group)
            nat.call("pc_transitions_sum", self.engine.h, self.corpus.c, _p(self.utt_logp), _p(self.pair_trans),
                     _p(self.tmax), _p(self.tsum), _stream())
            self._side_done = torch.cuda.Event()
            self._side_done.record(side)

    def reduce_statistics(self, group=None):
        """Accumulator reduction: the transition reductions (started by
        reduce_transitions_async, or here) joined into the current stream, then ONE SUM collective over
        the flat buffer [linear GMM statistics | transition sums] when `group` is a process group."""
        if getattr(self, "_side_done", None) is None:
            self.reduce_transitions_async(group)
        torch.cuda.current_stream().wait_event(self._side_done)
        self._side_done = None
        if getattr(self, "peer", None) is None:
            _dist.allreduce_flat_statistics(self.flat, group)

    reduce_transitions = reduce_statistics

    def estep(self, fix_code=0, group=None):
        """(K1 || log bands) -> K2 -> (K3 || transition reductions) -> accumulator reduction."""
        if getattr(self, "peer", None) is not None:
            self._peer_bind()  # this iteration's statistic set
        self.log_bands_async()
        self.score()
        self.forward_backward()
        self.reduce_transitions_async(group)
        if not (fix_code & 2):
            self.accumulate()
        else:
            self.acc.zero_()
        self.reduce_statistics(group)

    def mstep(self, c_covariance=1e-3, fix_code=0):
        m = self.model
        if getattr(self, "peer", None) is not None:
            nat.call("pc_update_params_peer", self.engine.h, m.mix, m.dim, _p(self.shift), _p(self.inv_scale),
                     float(c_covariance), int(fix_code), _p(m.mean), _p(m.var), _p(m.alpha), _p(m.transmat), _stream())
            self.acc, self.tsum, self.tmax = self.peer.views(2)  # the sums over the ranks
            return
        nat.call("pc_update_params", self.engine.h, m.n_units, m.mix, m.dim, _p(self.acc), _p(self.tmax),
                 _p(self.tsum), _p(self.shift), _p(self.inv_scale), float(c_covariance), int(fix_code),
                 _p(m.mean), _p(m.var), _p(m.alpha), _p(m.transmat), _stream())

    def em_iteration(self, c_covariance=1e-3, fix_code=0, group=None):
        self.estep(fix_code, group)
        self.mstep(c_covariance, fix_code)

    # views --------------------------------------------------------------------------------
    def linear_stats(self):
        """Accumulators in the ORIGINAL feature coordinates (fp64 numpy): occ [U,3,M] = sum gamma,
        sx [U,3,M,D] = sum gamma x, sxx [U,3,M,D] = sum gamma x^2 (SURVEY A.4's linear-equivalent
        form of Clustering.GMM's log-domain accumulators, Clustering.py:653-680)."""
        m = self.model
        a = self.acc.view(m.n_units, EMIT, m.mix, KA)
        D = m.dim
        occ = a[..., XS - 1]
        sx, sxx = a[..., :D], a[..., XS:XS + D]
        if self.shift is not None:
            sh, isc = self.shift[:D], self.inv_scale[:D]
            sx = sx / isc + sh * occ[..., None]
            sxx = sxx / (isc * isc) + 2 * sh * sx - sh * sh * occ[..., None]
        return occ.cpu().numpy(), sx.cpu().numpy(), sxx.cpu().numpy()

    def transition_accumulators(self):
        """(ksai_acc [U,3,5], gamma_acc [U,3]) in the reference's log domain (LHMM.py:84-85)."""
        acc = (self.tmax + torch.log(self.tsum)).cpu().numpy().reshape(-1, EMIT, 3)
        U = acc.shape[0]
        ksai = np.full((U, EMIT, STATES), -np.inf)
        for r in range(EMIT):
            ksai[:, r, r + 1] = acc[:, r, 0]
            ksai[:, r, r + 2] = acc[:, r, 1]
        return ksai, acc[:, :, 2].copy()


def viterbi(engine, corpus, b, log_self, log_next, utt_logpi=None, state_logpi=None, want_units=True):
    """pc_viterbi.  b: float32 or float64 emission buffer (corpus layout); log_self/log_next
    [U,5] fp64 device tensors computed on the host with numpy; logpi per utterance or per state."""
    path = engine.empty((corpus.total_frames,), torch.int32)
    units = engine.empty((corpus.total_frames,), torch.int32) if want_units else None
    score = engine.empty((corpus.n_utt,), torch.float64)
    b32 = b if b.dtype == torch.float32 else None
    b64 = b if b.dtype == torch.float64 else None
    nat.call("pc_viterbi", engine.h, corpus.c, _p(b32), _p(b64), _p(log_self), _p(log_next), _p(utt_logpi),
             _p(state_logpi), _p(path), _p(units), _p(score), _stream())
    return score, path, units


def host_log_bands(transmat, device):
    """np.log of the diagonal / super-diagonal of [U,5,5] transmat (host numpy: bit-exact with the
    reference's np.log(transmat), LHMM.py:340,574)."""
    tm = np.asarray(transmat, dtype=np.float64)
    with np.errstate(divide="ignore"):
        ls = np.log(np.diagonal(tm, axis1=1, axis2=2))
        ln = np.full_like(ls, -np.inf)
        ln[:, :-1] = np.log(np.diagonal(tm, offset=1, axis1=1, axis2=2))
    return (torch.as_tensor(np.ascontiguousarray(ls)).to(device), torch.as_tensor(np.ascontiguousarray(ln)).to(device))


def frame_moments_host(engine, frames, group=None):
    """Standardisation constants of a corpus held in host memory (pc_frame_moments_host): per-dimension
    mean and 1/std as fp64 numpy arrays.  With a process group the moments of all ranks are added first,
    so every rank maps its frames - and its accumulators, which live in the standardised space - the
    same way."""
    frames = np.ascontiguousarray(frames, dtype=np.float32)
    n, D = frames.shape
    s, q = np.zeros(D), np.zeros(D)
    nat.call("pc_frame_moments_host", engine.h, _p(frames), n, D, _p(s), _p(q), _stream())
    mom = np.concatenate([s, q, [float(n)]])
    if group is not None:
        t = torch.as_tensor(mom).to(engine.device)
        torch.distributed.all_reduce(t, group=group)
        mom = t.cpu().numpy()
    cnt = mom[-1]
    mu = mom[:D] / cnt
    sd = np.sqrt(np.maximum(mom[D:2 * D] / cnt - mu * mu, 0.0))
    return np.ascontiguousarray(mu), np.ascontiguousarray(1.0 / np.maximum(sd, 1e-12))


class HostReduceHook:
    """NCCL side of pc_em_iteration_host at world size > 1 (pc_set_reduce_hook): the exchange buffers
    are torch tensors; the library calls back between its launches, op 0 = all-reduce MAX of the
    transition maxima, op 1 = all-reduce SUM of [GMM statistics | transition sums], each queued on the
    stream it names."""

    def __init__(self, engine, n_units, n_gauss, group):
        self.engine, self.group = engine, group
        self.tmax = engine.empty((n_units, SLOTS), torch.float64)
        self.flat = engine.empty((n_gauss * KA + n_units * SLOTS,), torch.float64)
        self.error = None

        def _hook(user, op, stream):
            try:
                # a NULL stream is the (legacy) default stream; torch.cuda.ExternalStream(0) would NOT be that stream
                # - a zero pointer reads as "no pointer given" and yields a fresh pool stream
                ext = torch.cuda.ExternalStream(int(stream), device=engine.device) if stream else \
                    torch.cuda.default_stream(engine.device)
                with torch.cuda.stream(ext):
                    if op == 0:
                        torch.distributed.all_reduce(self.tmax, op=torch.distributed.ReduceOp.MAX, group=self.group)
                    else:
                        torch.distributed.all_reduce(self.flat, group=self.group)
                return 0
            except Exception as e:  # never let an exception cross the C boundary
                self.error = e
                return 1

        self._cb = nat.REDUCE_HOOK(_hook)  # keep the trampoline alive
        nat.call("pc_set_reduce_hook", engine.h, C.cast(self._cb, C.c_void_p), None, _p(self.tmax), _p(self.flat),
                 self.flat.numel())

    def remove(self):
        nat.call("pc_set_reduce_hook", self.engine.h, None, None, None, None, 0)


def em_iteration_host(engine, corpus, frames, mean, var, alpha, transmat, c_covariance=1e-3, fix_code=0,
                      shift=None, inv_scale=None):
    """pc_em_iteration_host: host numpy buffers in, parameters updated in place, returns sum logP.
    shift / inv_scale: the corpus' standardisation constants (frame_moments_host); None = computed
    from `frames` inside the call."""
    frames = np.ascontiguousarray(frames, dtype=np.float32)
    for a in (mean, var, alpha, transmat):
        if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous):
            raise TypeError("parameters must be C-contiguous float64 numpy arrays (updated in place)")
    if (shift is None) != (inv_scale is None):
        raise ValueError("shift and inv_scale go together")
    if shift is not None:
        shift = np.ascontiguousarray(shift, dtype=np.float64)
        inv_scale = np.ascontiguousarray(inv_scale, dtype=np.float64)
    U, _, M, D = mean.shape
    out = C.c_double(0.0)
    nat.call("pc_em_iteration_host", engine.h, corpus.c, _p(frames), D, U, M, _p(mean), _p(var), _p(alpha),
             _p(transmat), _p(shift), _p(inv_scale), float(c_covariance), int(fix_code), C.byref(out), _stream())
    return out.value


# ------------------------------------------------------------------------------------ K5 k-means
def kmeans_seed_points(x0, k, rnd):
    """Seeding of ClusterInitialization.kmeans(algorithm=1) (Clustering.py:975-1020):
    the draws come from Python's `random` (module or random.Random), so this part is host code by
    construction.  x0: metric coordinate (dimension 0, Q2) of the problem's points, fp64 numpy.
    Returns the k seed point indices (duplicates possible, Q9)."""
    n = len(x0)
    c0 = rnd.randint(0, n - 1)
    d = np.abs(x0[c0] - x0)
    dl = np.sqrt(d * d)  # (|d| ** 2) ** 0.5
    total = float(np.add.accumulate(dl)[-1])  # sequential fp64 sum in index order
    seeds = [c0]
    if total == 0.0:
        idx = rnd.sample(range(0, n), k - 1)
        if k > 2:
            raise AssertionError("all points coincide in dimension 0: the reference asserts for k > 2 "
                                 "(Clustering.py:1004-1008)")
        seeds.extend(idx)
        return seeds
    for _ in range(1, k):
        r = rnd.randint(0, int(total))
        run = np.subtract.accumulate(np.concatenate(([float(r)], dl)))[1:]  # r -= dl[i], sequentially
        hit = np.nonzero(run < 0)[0]
        if len(hit) == 0:
            raise AssertionError("k-means seeding found no point (Clustering.py:1008)")
        seeds.append(int(hit[0]))
    return seeds


def kmeans_run(engine, x, point_off, k, seeds, max_passes=1 << 40):
    """pc_kmeans_run + pc_kmeans_finish for a batch of independent problems.
    x: [P,D] fp64 cuda tensor (problems concatenated), point_off: host int64 [n_problems+1],
    seeds: int32 [n_problems,k] point indices inside each problem.
    Returns dict(owner, member_list, member_count, passes, moves, mean, var, alpha) of device tensors
    (member_list region of problem p starts at point_off[p] + p*k)."""
    point_off = np.ascontiguousarray(point_off, dtype=np.int64)
    n_problems = len(point_off) - 1
    P, D = x.shape
    if x.dtype != torch.float64:
        raise TypeError("k-means data must be float64 (the reference's precision)")
    if int(point_off[-1]) != P:
        raise ValueError("point_off[-1]=%d but x has %d rows" % (int(point_off[-1]), P))
    x = x.contiguous()
    dev = engine.device
    seeds_d = torch.as_tensor(np.ascontiguousarray(seeds, dtype=np.int32).reshape(n_problems, k)).to(dev)
    ws_bytes = int(nat.lib().pc_kmeans_workspace_bytes(n_problems, _p(point_off), int(k)))
    if ws_bytes < 0:
        raise ValueError("k-means: bad problem sizes or k outside [1,127]")
    ws = engine.empty((max(ws_bytes, 1),), torch.uint8)
    out = dict(owner=engine.empty((P,), torch.int32),
               member_list=engine.empty((P + n_problems * k,), torch.int32),
               member_count=engine.empty((n_problems, k), torch.int32),
               passes=engine.empty((n_problems,), torch.int32),
               moves=engine.empty((n_problems,), torch.int64),
               mean=engine.empty((n_problems, k, D), torch.float64),
               var=engine.empty((n_problems, k, D), torch.float64),
               alpha=engine.empty((n_problems, k), torch.float64))
    nat.call("pc_kmeans_run", engine.h, n_problems, _p(point_off), _p(x), D, int(k), _p(seeds_d), _p(ws),
             _p(out["owner"]), _p(out["member_list"]), _p(out["member_count"]), _p(out["passes"]),
             _p(out["moves"]), int(max_passes), _stream())
    nat.call("pc_kmeans_finish", engine.h, n_problems, _p(point_off), _p(x), D, int(k), _p(ws),
             _p(out["member_list"]), _p(out["member_count"]), _p(out["mean"]), _p(out["var"]),
             _p(out["alpha"]), _stream())
    out["workspace"] = ws  # keep alive until the stream has run
    return out


def gather_state_data(engine, data, key_off, group=None, max_points=None):
    """Per-(unit, state) data sets across ranks.  Every rank has grouped the frames of ITS utterances by
    state (`group_frames`); the reference pools the segments of all machines per unit on a shared file
    system before k-means (AcousticModel.py:532-561, Controller.py:47-106).  Here every rank contributes
    its first ceil(max_points / R) frames of each state and one all-gather hands every rank the pooled
    (capped) sets, rank-major.  Returns (data [P, D] fp64 device tensor, key_off host int64 [S + 1])."""
    key_off = np.asarray(key_off, dtype=np.int64)
    S = len(key_off) - 1
    R = 1 if group is None else torch.distributed.get_world_size(group)
    counts = np.diff(key_off)
    if max_points is None and R == 1:
        return data, key_off
    q = int(counts.max()) if max_points is None else int(-(-int(max_points) // R))
    D = data.shape[1]
    take = np.minimum(counts, q)
    if R == 1:
        sel = torch.cat([torch.arange(int(key_off[s]), int(key_off[s] + take[s]), device=data.device) for s in range(S)])
        return data[sel].contiguous(), np.concatenate([[0], np.cumsum(take)]).astype(np.int64)
    local = torch.zeros((S, q, D), dtype=data.dtype, device=data.device)
    for s in range(S):
        if take[s]:
            local[s, :int(take[s])] = data[int(key_off[s]):int(key_off[s] + take[s])]
    cnt = torch.as_tensor(take.astype(np.int64)).to(data.device)
    all_cnt = torch.empty((R, S), dtype=torch.int64, device=data.device)
    all_dat = torch.empty((R, S, q, D), dtype=data.dtype, device=data.device)
    torch.distributed.all_gather_into_tensor(all_cnt, cnt, group=group)
    torch.distributed.all_gather_into_tensor(all_dat, local, group=group)
    all_cnt = all_cnt.cpu().numpy()
    parts, off = [], [0]
    for s in range(S):
        n_s = 0
        for r in range(R):
            c = int(all_cnt[r, s])
            if c:
                parts.append(all_dat[r, s, :c])
                n_s += c
        off.append(off[-1] + n_s)
    return torch.cat(parts).contiguous(), np.asarray(off, dtype=np.int64)


def kmeans_states(engine, data, key_off, k, group=None, seed_rng=None):
    """k-means initialisation of every (unit, state) data set (ClusterInitialization.kmeans(algorithm=1),
    AcousticModel.py:553-554) with the STATES sharded over the ranks of `group` (state s -> rank s mod R:
    the problems are independent) and the parameters summed into place on every rank (disjoint
    supports: the all-reduce is an all-gather).  `data` / `key_off`: the pooled sets (`gather_state_data`),
    identical on every rank.  seed_rng(s) -> the `random.Random` that draws state s's seeds (the reference
    uses the module-level generator).  Returns dict(mean [S,k,D], var [S,k,D], alpha [S,k], passes [S])."""
    import random as _random

    key_off = np.asarray(key_off, dtype=np.int64)
    S = len(key_off) - 1
    D = data.shape[1]
    R, rank = 1, 0
    if group is not None:
        R, rank = torch.distributed.get_world_size(group), torch.distributed.get_rank(group)
    mine = [s for s in range(S) if s % R == rank and key_off[s + 1] - key_off[s] >= k]
    dev = engine.device
    mean = torch.zeros((S, k, D), dtype=torch.float64, device=dev)
    var = torch.zeros((S, k, D), dtype=torch.float64, device=dev)
    alpha = torch.zeros((S, k), dtype=torch.float64, device=dev)
    passes = torch.zeros((S,), dtype=torch.int64, device=dev)
    if mine:
        sub = torch.cat([data[int(key_off[s]):int(key_off[s + 1])] for s in mine]).contiguous()
        sub_off = np.concatenate([[0], np.cumsum([key_off[s + 1] - key_off[s] for s in mine])]).astype(np.int64)
        x0 = sub[:, 0].cpu().numpy()
        seeds = [kmeans_seed_points(np.ascontiguousarray(x0[sub_off[i]:sub_off[i + 1]]), k,
                                    seed_rng(s) if seed_rng is not None else _random.Random(s))
                 for i, s in enumerate(mine)]
        out = kmeans_run(engine, sub, sub_off, k, np.array(seeds, dtype=np.int32))
        idx = torch.as_tensor(np.asarray(mine, dtype=np.int64)).to(dev)
        mean[idx], var[idx], alpha[idx] = out["mean"], out["var"], out["alpha"]
        passes[idx] = out["passes"].to(torch.int64)
    if group is not None:
        for t in (mean, var, alpha, passes):
            torch.distributed.all_reduce(t, group=group)
    return dict(mean=mean, var=var, alpha=alpha, passes=passes, states=mine)


# ---- alignment post-processing ------------------------------------------------
def segment_keys(engine, corpus, path=None):
    """pc_segment_keys: per corpus frame the data set it joins, key = unit * 3 + emitting state, or -1.
    path=None: uniform segmentation (AcousticModel.__eq_segment mode 'e', AcousticModel.py:605-612);
    path = the composite-state path `viterbi` returned: the re-segmentation of multi_process_data
    (AcousticModel.py:750-764).  Either way each unit segment is cut in three (__get_gmmdata,
    :630-644).  Returns (frame_key int32 [F], utt_kept int32 [U]) device tensors."""
    F = int(corpus.frame_off[-1])
    key = engine.empty((F,), torch.int32)
    kept = engine.empty((len(corpus.frame_off) - 1,), torch.int32)
    if path is not None:
        if path.dtype != torch.int32 or path.numel() != F:
            raise ValueError("path must be the int32 [total_frames] tensor viterbi() returned")
        path = path.contiguous()
    nat.call("pc_segment_keys", engine.h, corpus.c, 0 if path is None else 1, _p(path), _p(key), _p(kept), _stream())
    return key, kept


def group_frames(engine, frame_key, n_keys, x=None):
    """pc_group_frames (+ pc_gather_rows when x is given): frames of one key made contiguous, in
    (utterance, time) order - the per-state data sets AcousticModel.__get_gmmdata builds.
    Returns dict(key_off: host int64 [n_keys+1], order: device int32 [F], data: x[order[:kept]] or None)."""
    frame_key = frame_key.contiguous()
    F = frame_key.numel()
    ws_bytes = int(nat.lib().pc_group_workspace_bytes(F, int(n_keys)))
    if ws_bytes < 0:
        raise ValueError("group_frames: %d frames / %d keys outside the supported range" % (F, n_keys))
    ws = engine.empty((ws_bytes,), torch.uint8)
    key_off = engine.empty((n_keys + 2,), torch.int64)
    order = engine.empty((max(F, 1),), torch.int32)
    nat.call("pc_group_frames", engine.h, _p(frame_key), F, int(n_keys), _p(ws), _p(key_off), _p(order), _stream())
    off = key_off.cpu().numpy()  # the caller sizes the per-state problems from it: one host sync
    kept = int(off[n_keys])
    data = None
    if x is not None:
        if x.dim() != 2 or x.shape[0] != F:
            raise ValueError("x must hold one row per corpus frame")
        x = x.contiguous()
        row_bytes = x.shape[1] * x.element_size()
        data = engine.empty((kept, x.shape[1]), x.dtype)
        nat.call("pc_gather_rows", engine.h, _p(order), kept, row_bytes, _p(x), _p(data), _stream())
    return dict(key_off=off[:n_keys + 1].copy(), order=order[:F], data=data)

