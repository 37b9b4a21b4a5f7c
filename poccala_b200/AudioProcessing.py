"""Mirror of the reference's feature front end (StatisticalModel/AudioProcessing.py): `AudioProcessing.MFCC` and
`AudioProcessing.VAD` with the reference's names and arguments, computing on the device (csrc/mfcc.cu) what
`AcousticModel.__load_audio` needs (AcousticModel.py:463-477): MFCC(13).init_audio(path=...) -> mfcc(nfft=512,
d1=..., d2=...) -> VAD().init_mfcc(m) -> mfcc().  SURVEY section 8 f4.

The arithmetic is the reference's as written (the kernel header lists the differences from a
textbook MFCC).  Recording / playback (pyaudio) and plotting (pylab) are outside this path: `show_pic` is accepted
and ignored, `RecordAudio` is not provided.
"""
from __future__ import annotations

import math
import wave

import numpy as np
import torch

from . import _native as nat
from ._native import _p
from .engine import _stream
from .runtime import get_engine


class AudioProcessing(object):
    class MFCC(object):
        def __init__(self, vec_num=13):
            self.__wav = None
            self.__wdata = None
            self.__params = None
            self.__vec_num = vec_num

        @property
        def data(self):
            return self.__wdata

        @property
        def wav(self):
            return self.__wav

        @property
        def params(self):
            """(nchannels, sampwidth, framerate, nframes, comptype, compname) of the wave file."""
            return self.__params

        def init_audio(self, wav=None, path=None, show_pic=False):
            """AudioProcessing.py:146-183: 16-bit samples of a wave file (object or path); two channels are merged
            by taking the larger sample; samples equal to zero are REMOVED (:176)."""
            if wav is None:
                if path is None:
                    raise ValueError("init_audio: neither a wave object nor a path")
                wav = wave.open(path, "rb")
            self.__wav = wav
            self.__params = wav.getparams()
            data = np.frombuffer(wav.readframes(wav.getnframes()), dtype=np.short).copy()
            if wav.getnchannels() == 2:
                data = data.reshape(-1, 2).T
                data = np.where(data[0] < data[1], data[1], data[0])
            self.__wdata = np.delete(data, np.where(data == 0))

        def set_signal(self, samples, framerate):
            """Samples that do not come from a wave file (tests, streaming): int16 array + sample rate."""
            self.__wdata = np.asarray(samples)
            self.__params = (1, 2, int(framerate), len(self.__wdata), "NONE", "not compressed")

        @staticmethod
        def frame_geometry(samplenum, framerate, sampletime=0.025, overlap=0.5):
            """frame_blocking's arithmetic (AudioProcessing.py:215-219): (framesize, step, framenum)."""
            framesize = int(framerate * sampletime)
            step = int(framesize * overlap)
            framenum = 1 + math.ceil((samplenum - framesize) / step)
            return framesize, step, framenum

        @staticmethod
        def filter_responses(samplerate, nfft=512, low_hz=0., high_hz=None, filterbanks=26):
            """The response matrix of mel_filter_bank (AudioProcessing.py:302-343), [filterbanks][nfft // 2 + 1], with
            the reference's own expressions (natural-log mel scale, floor((nfft + 1) / rate * hz) bins, both flanks of a
            triangle rising)."""
            top = high_hz or samplerate / 2
            to_mel = lambda f: 2595 * math.log(1 + f / 700, math.e)  # noqa: E731  (natural log, as the reference has it)
            edges_hz = 700 * (np.exp(np.linspace(to_mel(low_hz), to_mel(top), filterbanks + 2) / 2595) - 1)
            edge = np.floor((nfft + 1) / samplerate * edges_hz)  # float bin positions, filterbanks + 2 of them
            lo = edge.astype(np.int64)
            out = np.zeros((filterbanks, nfft // 2 + 1))
            for m in range(filterbanks):
                # two ramps, each starting at 0 on its left edge: [lo_m, lo_m+1) and [lo_m+1, lo_m+2)
                for a, b in ((m, m + 1), (m + 1, m + 2)):
                    cols = np.arange(lo[a], lo[b])
                    out[m, cols] = (cols - lo[a]) / (edge[b] - edge[a])
            return out

        def mfcc(self, sampletime=0.025, overlap=0.5, nfft=512, cal_energy=True, d1=False, d2=False):
            """AudioProcessing.py:416-448 -> [frames, vec_num * (1 + d1 + d2)] fp64 numpy (d2 only with d1, as there)."""
            if self.__wdata is None:
                raise ValueError("mfcc: no audio loaded (init_audio)")
            eng = get_engine()
            rate = self.__params[2]
            n = len(self.__wdata)
            framesize, step, framenum = self.frame_geometry(n, rate, sampletime, overlap)
            if framenum < 2:
                raise ZeroDivisionError("a single frame: the reference's window divides by (frames - 1) = 0 "
                                        "(AudioProcessing.py:245)")
            n_delta = (1 if d1 else 0) + (1 if d1 and d2 else 0)
            sig = torch.as_tensor(np.ascontiguousarray(self.__wdata, dtype=np.float64)).to(eng.device)
            fb = torch.as_tensor(self.filter_responses(rate, nfft=nfft)).to(eng.device)
            out = torch.empty((framenum, self.__vec_num * (1 + n_delta)), dtype=torch.float64, device=eng.device)
            nat.call("pc_mfcc", eng.h, _p(sig), n, framesize, step, framenum, int(nfft), _p(fb), fb.shape[0], self.__vec_num,
                     1 if cal_energy else 0, n_delta, _p(out), _stream())
            return out.cpu().numpy()

    class VAD(object):
        def __init__(self, simple_size=16):
            self.__mfcc = None
            self.__simple_size = simple_size

        def init_mfcc(self, mfcc):
            self.__mfcc = np.ascontiguousarray(mfcc, dtype=np.float64)

        def _distances(self, alpha=0.5, beta=0.93):
            eng = get_engine()
            m = torch.as_tensor(self.__mfcc).to(eng.device)
            T, D = m.shape
            dist = torch.empty((T,), dtype=torch.float64, device=eng.device)
            osf = torch.empty_like(dist)
            nat.call("pc_vad_distance", eng.h, _p(m), T, D, self.__simple_size, float(alpha), float(beta), _p(dist), _p(osf),
                     _stream())
            return dist.cpu().numpy(), osf.cpu().numpy()

        def mel_distance(self, alpha=0.5):
            """AudioProcessing.py:462-478: distance of every frame to the running noise estimate."""
            return self._distances(alpha=alpha)[0]

        def osf(self, mel_distance=None, beta=0.93):
            """AudioProcessing.py:480-506 (computed from the frames held; `mel_distance` is accepted for the
            reference's signature)."""
            return self._distances(beta=beta)[1]

        def detect(self, mel_distance, show_pic=False):
            """AudioProcessing.py:508-534: frames whose smoothed distance exceeds d_mid (max - min) / max, d_mid the
            distance of frame simple_size / 2 (the sorted copy the reference makes is not used there either)."""
            mel_distance = np.asarray(mel_distance, dtype=np.float64)
            d_mid = mel_distance[int(self.__simple_size / 2)]
            max_distance, min_distance = mel_distance.max(), mel_distance.min()
            beta = d_mid * (max_distance - min_distance) / max_distance
            check = mel_distance - np.ones_like(mel_distance) * beta
            return self.__mfcc[np.where(check > 0.)]

        def mfcc(self, show_pic=False):
            """AudioProcessing.py:536-541: the frames that survive the detector."""
            return self.detect(self.osf(self.mel_distance()), show_pic=show_pic)
