"""Oracle vs the reference's own (few) pins for the mel path, independent implementations, and the
committed goldens.  CPU only."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err
from oracle import dsp


def test_reference_nil_cases():
    # Tests/IncrementalMelSpectrogramTests.swift:7-17 (nMels 128): 1- and 2-sample first chunk -> nil
    m = dsp.IncrementalMelSpectrogram(sample_rate=16000, n_fft=400, hop_length=160, n_mels=128)
    assert m.process([0.1]) is None
    m = dsp.IncrementalMelSpectrogram(sample_rate=16000, n_fft=400, hop_length=160, n_mels=128)
    assert m.process([0.1, -0.2]) is None
    assert dsp.IncrementalMelSpectrogram().process([]) is None


def test_whisper_feature_shape_for_5s_zeros():
    # Tests/MLXAudioSTTTests.swift:4416-4422
    f = dsp.whisper_encoder_features(np.zeros(16000 * 5, dtype=np.float32), 80)
    assert f.shape == (1, 3000, 80)


def test_hann_windows_closed_form():
    w = dsp.hanning_window(400)
    assert w[0] == 0 and abs(w[-1]) < 1e-6 and abs(w[199] - w[200]) < 1e-6      # symmetric (DSP.swift:17)
    wp = dsp.periodic_hann_window(400)
    assert wp[0] == 0 and abs(wp[200] - 1) < 1e-6                                   # periodic (WhisperAudio.swift:43)
    assert np.allclose(wp, torch.hann_window(400, periodic=True).numpy(), atol=1e-6)
    assert np.allclose(w, torch.hann_window(400, periodic=False).numpy(), atol=1e-6)


def test_whisper_matches_hf_feature_extractor():
    from transformers import WhisperFeatureExtractor
    x = dsp.synth_audio(160000, 0)
    ref = WhisperFeatureExtractor()(x, sampling_rate=16000, return_tensors="np").input_features[0]   # [80, 3000]
    ours = dsp.whisper_encoder_features(x, 80)[0]
    assert np.abs(ref.T - ours).max() < 1e-4


def test_core_power_spectrum_matches_torch_stft():
    x = dsp.synth_audio(16000, 1)
    padded = dsp.reflect_pad_core(x, 200)
    st = torch.stft(torch.from_numpy(x).double(), 400, 160, window=torch.from_numpy(dsp.hanning_window(400)).double(),
                    center=True, pad_mode="reflect", return_complex=True)
    power = (st.abs() ** 2).numpy().T                                                # [F, 201]
    F = 1 + (len(padded) - 400) // 160
    fr = dsp._frames(padded.astype(np.float64), F, 400, 160) * dsp.hanning_window(400).astype(np.float64)
    mine = np.abs(np.fft.rfft(fr, axis=1)) ** 2
    assert power.shape == mine.shape and rel_err(mine, power) < 1e-9


def test_streaming_equals_offline_single_chunk():
    # reference pattern: streaming == offline (VoxtralRealtimeStreamingFrontEndTests) -- holds for the
    # incremental class when the global max falls in the first chunk (SURVEY 8c trap 3)
    x = dsp.synth_audio(160000, 0)
    m = dsp.IncrementalMelSpectrogram(16000, 400, 160, 80)
    a, b = m.process(x), m.flush()
    off = dsp.compute_mel_spectrogram(x, 16000, 400, 160, 80)
    assert a.shape == (999, 80) and b.shape == (2, 80) and m.total_frames == 1001
    assert np.abs(off[:999] - a).max() < 1e-9
    assert off.shape[0] == 1001


def test_irregular_chunking_frame_count_and_monotone_max():
    x = dsp.synth_audio(48000, 2)
    m = dsp.IncrementalMelSpectrogram(16000, 400, 160, 80)
    rng = np.random.default_rng(0)
    cuts = np.sort(rng.choice(np.arange(1, 48000), size=17, replace=False)).tolist()
    cuts = [0] + cuts + [48000]
    n, prev = 0, -np.inf
    for i in range(len(cuts) - 1):
        o = m.process(x[cuts[i]:cuts[i + 1]])
        assert m.running_max >= prev
        prev = m.running_max
        if o is not None:
            n += o.shape[0]
    o = m.flush()
    n += o.shape[0]
    assert n == m.total_frames == 1 + 48000 // 160


def test_goldens():
    g = np.load(GOLDEN / "mel.npz")
    x = dsp.synth_audio(160000, 0)
    m = dsp.IncrementalMelSpectrogram(16000, 400, 160, 80)
    full = np.concatenate([m.process(x), m.flush()])
    assert tuple(g["inc_shape"]) == full.shape
    assert np.abs(full[:4] - g["inc_first"]).max() < 1e-6 and np.abs(full[-3:] - g["inc_last"]).max() < 1e-6
    w = dsp.whisper_encoder_features(x, 80)[0]
    assert np.abs(w[[0, 1, 999, 1000, 2999]] - g["whisper_rows"]).max() < 1e-6


def test_mel_filterbank_matches_torchaudio():
    """melFilters (DSP.swift:76-168) against torchaudio.functional.melscale_fbanks -- an independent implementation -- for the two
    configurations on the path: HTK scale + Slaney norm (IncrementalMelSpectrogram / computeMelSpectrogram) and Slaney + Slaney
    (Whisper).  The only structural difference is the reference's inclusive upper edge (`<=`, :149): at most one extra tiny entry."""
    torchaudio = pytest.importorskip("torchaudio")
    import warnings
    for scale in ("htk", "slaney"):
        for n_mels in (80, 128):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ref = torchaudio.functional.melscale_fbanks(n_freqs=201, f_min=0.0, f_max=8000.0, n_mels=n_mels, sample_rate=16000, norm="slaney",
                                                            mel_scale=scale).double().numpy()
            o = dsp.mel_filters(16000, 400, n_mels, norm="slaney", mel_scale=scale)
            assert o.shape == ref.shape == (201, n_mels)
            assert np.abs(o - ref).max() < 5e-7 * max(1.0, np.abs(ref).max() / 0.03)
            extra = (o != 0) & (ref == 0)
            assert extra.sum() <= 1 and not ((o == 0) & (ref != 0)).any()


def test_offline_log_mel_matches_torchaudio_pipeline():
    """computeMelSpectrogram (DSP.swift:181-273: reflect pad both sides, symmetric Hann, |rfft|^2, HTK/Slaney filterbank, log10, global
    max - 8 clamp, (x + 4) / 4) against torchaudio's MelSpectrogram (independent STFT and filterbank) with the same post-processing."""
    torchaudio = pytest.importorskip("torchaudio")
    x = dsp.synth_audio(160000, 0)
    o = dsp.compute_mel_spectrogram(x, 16000, 400, 160, 80)
    win = torch.from_numpy(dsp.hanning_window(400)).double()
    T = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=400, hop_length=160, n_mels=80, f_min=0.0, f_max=8000.0, power=2.0, center=True,
                                             pad_mode="reflect", norm="slaney", mel_scale="htk", window_fn=lambda n: win).double()
    m = T(torch.from_numpy(x).double()).numpy().T
    assert m.shape == o.shape == (1001, 80)                        # SURVEY config 1: 999 + 2 frames
    lg = np.log10(np.maximum(m, 1e-10))
    lg = (np.maximum(lg, lg.max() - 8.0) + 4.0) / 4.0
    assert np.abs(lg - o).max() < 1e-4


def test_streaming_first_chunk_matches_torchaudio_frames():
    """IncrementalMelSpectrogram.process on ONE chunk (IncrementalMelSpectrogram.swift:68-147): the reflect prefix of the first chunk makes
    its frames the centre-padded STFT frames, so they must equal torchaudio's MelSpectrogram frames (independent STFT + filterbank) under
    the same log / running-max clamp / (x + 4) / 4, with the maximum taken over the frames that chunk produced."""
    torchaudio = pytest.importorskip("torchaudio")
    x = dsp.synth_audio(160000, 0)
    m = dsp.IncrementalMelSpectrogram(16000, 400, 160, 80)
    a = m.process(x)
    assert a.shape == (999, 80)                                       # SURVEY config 1: 999 frames from process, 2 more from flush
    win = torch.from_numpy(dsp.hanning_window(400)).double()
    T = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=400, hop_length=160, n_mels=80, f_min=0.0, f_max=8000.0, power=2.0, center=True,
                                             pad_mode="reflect", norm="slaney", mel_scale="htk", window_fn=lambda n: win).double()
    lg = np.log10(np.maximum(T(torch.from_numpy(x).double()).numpy().T[:999], 1e-10))
    ref = (np.maximum(lg, lg.max() - 8.0) + 4.0) / 4.0
    assert np.abs(a - ref).max() < 1e-4
