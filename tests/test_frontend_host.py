"""§8 f4, host half of the front end: wave decoding (AudioProcessing.py:146-183), frame geometry (:215-219) and the
mel filter responses (:302-343) against the golden run of the executed reference (tests/golden/mfcc.npz).  The device
half is tests/test_gpu_frontend.py."""
import os
import wave

import numpy as np
import pytest

from tests.helpers import load_golden

torch = pytest.importorskip("torch")


def _write_wav(path, pcm, rate, channels):
    w = wave.open(str(path), "wb")
    w.setnchannels(int(channels)); w.setsampwidth(2); w.setframerate(int(rate)); w.writeframes(np.asarray(pcm, np.int16).tobytes())
    w.close()


def test_wave_decoding_and_geometry(tmp_path):
    from poccala_b200.AudioProcessing import AudioProcessing

    g = load_golden("mfcc.npz")
    for c in range(int(g["n"])):
        path = tmp_path / ("x%d.wav" % c)
        _write_wav(path, g[f"m{c}_pcm"], g[f"m{c}_rate"], g[f"m{c}_channels"])
        m = AudioProcessing.MFCC(13)
        m.init_audio(path=str(path))
        assert m.data.dtype == np.int16 and np.array_equal(m.data, g[f"m{c}_data"])  # stereo merge, zeros removed
        assert m.params[2] == int(g[f"m{c}_rate"])
        framesize, step, framenum = m.frame_geometry(len(m.data), m.params[2])
        assert framenum == g[f"m{c}_mfcc"].shape[0] and framesize == int(m.params[2] * 0.025) and step == framesize // 2


def test_filter_responses_shape_and_quirk():
    """Both flanks of a triangle rise (AudioProcessing.py:322-326): the response drops to 0 at the centre bin and
    climbs again, and every filter covers its three mel-spaced bins."""
    from poccala_b200.AudioProcessing import AudioProcessing

    r = AudioProcessing.MFCC.filter_responses(16000, nfft=512)
    assert r.shape == (26, 257) and (r >= 0).all() and (r < 1).all()
    nz = [np.flatnonzero(row) for row in r]
    assert all(len(k) > 0 for k in nz)
    assert all(nz[i][0] <= nz[i + 1][0] for i in range(25))
    # rising towards the centre, zero AT the centre bin, rising again after it
    row = r[20]
    centre = nz[20][np.argmax(np.diff(row[nz[20]]) < 0) ] + 1 if (np.diff(row[nz[20]]) < 0).any() else None
    assert centre is not None and row[centre] == 0.0
