"""CUDA mel front-end (through the C ABI) vs the oracle, the committed goldens and size-independent
properties.  Tolerance: 1e-3 relative (north_star) -- measured here as max |diff| / max |ref| on the
(x+4)/4-scaled log-mel, which is O(1)."""
import numpy as np
import pytest

from conftest import GOLDEN, max_rel_to_peak
from oracle import dsp

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _stream(cls, x, cuts, **kw):
    m = cls(16000, 400, 160, **kw)
    outs = []
    for i in range(len(cuts) - 1):
        o = m.process(x[cuts[i]:cuts[i + 1]])
        outs.append(o)
    outs.append(m.flush())
    return m, outs


def test_reference_nil_cases(b2a):
    # Tests/IncrementalMelSpectrogramTests.swift:7-17
    assert b2a.IncrementalMelSpectrogram(16000, 400, 160, 128).process([0.1]) is None
    assert b2a.IncrementalMelSpectrogram(16000, 400, 160, 128).process([0.1, -0.2]) is None
    m = b2a.IncrementalMelSpectrogram(16000, 400, 160, 128)
    assert m.process([]) is None and m.flush() is None and m.total_frames == 0


def test_config1_single_chunk_vs_oracle_and_golden(b2a):
    x = dsp.synth_audio(160000, 0)
    m, (a, b) = _stream(b2a.IncrementalMelSpectrogram, x, [0, 160000], n_mels=80)
    mo, (oa, ob) = _stream(dsp.IncrementalMelSpectrogram, x, [0, 160000], n_mels=80)
    assert a.shape == (999, 80) and b.shape == (2, 80) and m.total_frames == 1001
    assert max_rel_to_peak(a, oa) < TOL and max_rel_to_peak(b, ob) < TOL
    g = np.load(GOLDEN / "mel.npz")
    full = np.concatenate([a, b])
    assert max_rel_to_peak(full[:4], g["inc_first"]) < TOL and max_rel_to_peak(full[-3:], g["inc_last"]) < TOL
    assert abs(full.mean() - g["inc_stats"][0]) < 1e-3 and abs(full.max() - g["inc_stats"][3]) < 1e-3


@pytest.mark.parametrize("n_mels", [80, 128])
def test_irregular_chunks_chunk_for_chunk(b2a, n_mels):
    x = dsp.synth_audio(160000, 0)
    cuts = [0, 1, 3, 150, 700, 5000, 5160, 40000, 160000]
    m, outs = _stream(b2a.IncrementalMelSpectrogram, x, cuts, n_mels=n_mels)
    mo, oo = _stream(dsp.IncrementalMelSpectrogram, x, cuts, n_mels=n_mels)
    assert m.total_frames == mo.total_frames
    for a, o in zip(outs, oo):
        assert (a is None) == (o is None)
        if a is not None:
            assert a.shape == o.shape and max_rel_to_peak(a, o) < TOL


def test_short_first_chunk_repeated_prefix_and_reset(b2a):
    # first chunk shorter than the 200-sample reflect prefix (IncrementalMelSpectrogram.swift:84-93)
    x = dsp.synth_audio(3000, 4)
    cuts = [0, 37, 120, 1000, 3000]
    m, outs = _stream(b2a.IncrementalMelSpectrogram, x, cuts, n_mels=80)
    mo, oo = _stream(dsp.IncrementalMelSpectrogram, x, cuts, n_mels=80)
    for a, o in zip(outs, oo):
        assert (a is None) == (o is None)
        if a is not None:
            assert max_rel_to_peak(a, o) < TOL
    m.reset()
    assert m.total_frames == 0
    a = m.process(x)
    o = dsp.IncrementalMelSpectrogram(16000, 400, 160, 80).process(x)
    assert max_rel_to_peak(a, o) < TOL


def test_quiet_and_loud_signals(b2a):
    # dynamic range: pure tone (most bins at the max-8 clamp), silence (all at the 1e-10 floor)
    t = np.arange(32000) / 16000.0
    for x in (np.sin(2 * np.pi * 1000 * t).astype(np.float32), np.zeros(32000, np.float32),
              (1e-4 * np.sin(2 * np.pi * 440 * t)).astype(np.float32)):
        a = b2a.IncrementalMelSpectrogram(16000, 400, 160, 80).process(x)
        o = dsp.IncrementalMelSpectrogram(16000, 400, 160, 80).process(x)
        assert np.abs(a - o).max() < 2e-3


def test_whisper_features_batch(b2a):
    xs = np.stack([dsp.synth_audio(480000, i) for i in range(2)])
    lm = b2a.LogMel("whisper", n_mels=80)
    out = lm(xs)
    assert out.shape == (2, 3000, 80)
    for i in range(2):
        o = dsp.whisper_encoder_features(xs[i], 80)[0]
        assert max_rel_to_peak(out[i], o) < TOL
    g = np.load(GOLDEN / "mel.npz")
    assert max_rel_to_peak(out[0][[0, 1, 999, 1000, 2999]], dsp.whisper_encoder_features(xs[0])[0][[0, 1, 999, 1000, 2999]]) < TOL
    # pad-to-30s: a 10 s clip equals the same clip zero-padded by the caller
    short = dsp.synth_audio(160000, 0)
    a = lm(short[None])
    assert a.shape == (1, 3000, 80) and max_rel_to_peak(a[0][[0, 1, 999, 1000, 2999]], g["whisper_rows"]) < TOL
    # the reference's own shape pin: 5 s of zeros -> [1, 3000, 80] (Tests/MLXAudioSTTTests.swift:4416-4422)
    assert b2a.whisper_encoder_features(np.zeros(80000, np.float32)).shape == (1, 3000, 80)


def test_core_offline_and_stream_equals_offline(b2a):
    x = dsp.synth_audio(160000, 0)
    off = b2a.compute_mel_spectrogram(x, 16000, 400, 160, 80)
    o = dsp.compute_mel_spectrogram(x, 16000, 400, 160, 80)
    assert off.shape == (1001, 80) and max_rel_to_peak(off, o) < TOL
    m = b2a.IncrementalMelSpectrogram(16000, 400, 160, 80)
    a = m.process(x)
    assert np.abs(a - off[:999]).max() < 1e-5     # streaming == offline when the max falls in chunk 1


def test_invalid_inputs_are_errors(b2a):
    with pytest.raises(b2a.AudioGenerationError) as e:
        b2a.IncrementalMelSpectrogram(16000, 512, 160, 80)
    assert e.value.case == "invalidInput"
    with pytest.raises(b2a.AudioGenerationError):
        b2a.LogMel("core")(np.zeros((1, 100), np.float32))
