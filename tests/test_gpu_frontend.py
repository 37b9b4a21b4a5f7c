"""§8 f4, device half of the front end: AudioProcessing.MFCC.mfcc and AudioProcessing.VAD (AudioProcessing.py:416-542)
through pc_mfcc / pc_vad_distance, and AcousticModel.load_audio (AcousticModel.py:463-477), against the golden run of
the executed reference.  fp64 on both sides; the FFT algorithms differ (radix-2 here, pocketfft there), hence 1e-9."""
import wave

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from tests.helpers import load_golden  # noqa: E402

pytestmark = pytest.mark.gpu


def _write_wav(path, pcm, rate, channels):
    w = wave.open(str(path), "wb")
    w.setnchannels(int(channels)); w.setsampwidth(2); w.setframerate(int(rate)); w.writeframes(np.asarray(pcm, np.int16).tobytes())
    w.close()


def _close(a, b, tol=1e-9):
    return np.all(np.abs(a - b) <= tol * np.maximum(np.abs(b), 1.0))


def test_mfcc_and_vad_match_executed_reference(tmp_path):
    from poccala_b200.AcousticModel import AcousticModel
    from poccala_b200.AudioProcessing import AudioProcessing

    g = load_golden("mfcc.npz")
    for c in range(int(g["n"])):
        path = tmp_path / ("x%d.wav" % c)
        _write_wav(path, g[f"m{c}_pcm"], g[f"m{c}_rate"], g[f"m{c}_channels"])
        m = AudioProcessing.MFCC(13)
        m.init_audio(path=str(path))
        feat = m.mfcc(nfft=512, d1=True, d2=True)
        assert feat.shape == g[f"m{c}_mfcc"].shape
        assert _close(feat, g[f"m{c}_mfcc"]), np.abs(feat - g[f"m{c}_mfcc"]).max()
        static = m.mfcc(nfft=512, cal_energy=False)
        assert static.shape == (feat.shape[0], 13) and _close(static, g[f"m{c}_static"])
        assert _close(static[:, 1:], feat[:, 1:13])  # only coefficient 0 is replaced by the log energy
        v = AudioProcessing.VAD()
        v.init_mfcc(feat)
        dist = v.mel_distance()
        osf = v.osf(dist)
        assert _close(dist, g[f"m{c}_dist"]) and _close(osf, g[f"m{c}_osf"])
        kept = v.mfcc()
        assert kept.shape == g[f"m{c}_kept"].shape and _close(kept, g[f"m{c}_kept"])
        # the surface AcousticModel uses
        am = AcousticModel(None, "T", state_num=5, mix_level=2)
        assert _close(am.load_audio(str(path)), g[f"m{c}_kept"])
    with pytest.raises(ZeroDivisionError):
        one = AudioProcessing.MFCC(13)
        one.set_signal(np.arange(1, 300, dtype=np.int16), 16000)
        one.mfcc()


@pytest.mark.parametrize("rate,nfft,vec,d1,d2,n", [(44100, 512, 13, True, True, 30000), (16000, 1024, 12, True, False, 9000),
                                                   (8000, 256, 13, False, False, 2600), (22050, 2048, 20, True, True, 12000)])
def test_mfcc_other_geometries_match_the_oracle(rate, nfft, vec, d1, d2, n):
    """Frames longer than the FFT (44.1 kHz: 1 102 samples cut to 512, AudioProcessing.py:262), other FFT lengths and
    coefficient counts, with and without deltas, against the oracle restatement (itself pinned to the executed reference
    by tests/test_oracle_golden.py)."""
    from oracle import fast
    from poccala_b200.AudioProcessing import AudioProcessing

    rng = np.random.default_rng(rate + nfft)
    t = np.arange(n) / rate
    sig = (2500 * np.sin(2 * np.pi * 310 * t) * (t > 0.2 * t[-1]) + 150 * rng.normal(size=n)).astype(np.int16)
    sig = sig[sig != 0]
    m = AudioProcessing.MFCC(vec)
    m.set_signal(sig, rate)
    got = m.mfcc(nfft=nfft, d1=d1, d2=d2)
    want = fast.mfcc_features(sig, rate, vec_num=vec, nfft=nfft, d1=d1, d2=d2)
    assert got.shape == want.shape
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), fin)
    assert _close(got[fin], want[fin]), np.abs(got[fin] - want[fin]).max()
    if got.shape[0] >= 40:
        v = AudioProcessing.VAD()
        v.init_mfcc(np.where(fin, got, 0.0))
        d, o, k = fast.vad_filter(np.where(fin, want, 0.0))
        assert _close(v.mel_distance(), d) and _close(v.osf(None), o)
        kept = v.mfcc()
        assert kept.shape == k.shape and _close(kept, k)
