"""Host logic of the streaming STT front end (row N3): StreamingEncoder windowing / overlap / cache and the decode cadence of
feedAudio, with the oracle's incremental mel and a stand-in encoder.  CPU only (tests-only use of the oracle)."""
import numpy as np
import pytest

from oracle import dsp


@pytest.fixture(scope="module")
def streaming():
    import importlib
    return importlib.import_module("mlx_audio_swift_b200.streaming")


class FakeEncoder:
    """encodeSingleWindow stand-in: 8x frame pooling (mean) so that outputs identify their input frames."""
    n_window_infer = 800

    def __init__(self):
        self.calls = []

    def encode_single_window(self, frames):
        frames = np.asarray(frames)
        self.calls.append(frames.shape[0])
        n = -(-frames.shape[0] // 8)
        pad = np.concatenate([frames, np.zeros((n * 8 - frames.shape[0], frames.shape[1]), frames.dtype)])
        return pad.reshape(n, 8, -1).mean(1)


def test_config_defaults_and_overlap(streaming):
    c = streaming.StreamingConfig()
    assert (c.decode_interval_seconds, c.boundary_decode_interval_seconds, c.boundary_boost_seconds, c.encoder_window_overlap_seconds) == (1.0, 0.2, 1.0, 1.0)
    assert (c.max_cached_windows, c.max_tokens_per_pass, c.min_agreement_passes, c.boundary_min_agreement_passes, c.max_decode_windows) == (60, 512, 2, 3, 1)
    assert c.finalize_completed_windows and c.language == "English" and c.delay_ms == 480
    assert streaming.StreamingConfig(delay_preset="realtime").delay_ms == 200 and streaming.StreamingConfig(delay_preset=750).delay_ms == 750
    assert c.overlap_frames(16000) == 100                          # 1 s of 10 ms hops (StreamingInferenceSession.swift:982)


def test_windows_overlap_and_cache(streaming):
    enc = FakeEncoder()
    se = streaming.StreamingEncoder(enc, max_cached_windows=2, overlap_frames=100)
    assert (se.window_size, se.window_stride) == (800, 700)
    frames = np.arange(2500 * 4, dtype=np.float32).reshape(2500, 4)
    assert se.feed(frames[:799]) == 0 and se.has_pending_frames and se.encoded_window_count == 0
    assert se.encode_pending().shape == (100, 4) and se.has_pending_frames          # early feedback does not consume
    assert se.feed(frames[799:1600]) == 2                          # windows [0, 800) and [700, 1500)
    assert enc.calls[-2:] == [800, 800] and se.encoded_window_count == 2
    new = se.drain_newly_encoded_windows()
    assert len(new) == 2 and se.drain_newly_encoded_windows() == []
    assert np.array_equal(new[1], enc.encode_single_window(frames[700:1500]))
    assert se.feed(frames[1600:2500]) == 1                         # window [1400, 2200); frames [2100, 2500) stay pending (the overlap is kept)
    assert se.encoded_window_count == 3 and se.total_cached_tokens == 200          # cache capped at 2 windows, count monotonic
    assert np.array_equal(se.get_cached_encoder_output(), np.concatenate([enc.encode_single_window(frames[700:1500]), enc.encode_single_window(frames[1400:2200])]))
    assert se.get_cached_encoder_output(from_window=1).shape == (100, 4) and se.get_cached_encoder_output(from_window=2) is None
    full = se.get_full_encoder_output()
    assert full.shape == (200 + 50, 4)                             # + 400 / 8 tokens of the pending part
    assert se.flush_partial() == 1 and not se.has_pending_frames and se.flush_partial() == 0
    assert se.encoded_window_count == 3                            # a flushed partial is cached but not counted (:101-116)
    se.reset()
    assert se.get_full_encoder_output() is None and se.encoded_window_count == 0
    # overlap is clamped to window - 1 and the stride to >= 1 (:47-51)
    assert streaming.StreamingEncoder(FakeEncoder(), overlap_frames=5000).window_stride == 1
    assert streaming.StreamingEncoder(FakeEncoder(), overlap_frames=-3).window_stride == 800


def test_feed_audio_cadence_with_the_oracle_mel(streaming):
    now = [0.0]
    mel = dsp.IncrementalMelSpectrogram(sample_rate=16000, n_fft=400, hop_length=160, n_mels=128)
    enc = FakeEncoder()
    cfg = streaming.StreamingConfig()
    fe = streaming.StreamingFrontEnd(mel, streaming.StreamingEncoder(enc, cfg.max_cached_windows, cfg.overlap_frames(16000)), cfg, clock=lambda: now[0])
    audio = dsp.synth_audio(16000 * 10, 3)
    assert fe.feed_audio(audio[:100]) is False                     # fewer samples than one frame: the mel returns nil
    assert fe.feed_audio(audio[100:16000]) is True                 # first content: decode immediately (no previous decode)
    assert fe.is_decoding and not fe.last_pass_is_boundary_finalize and fe.min_agreement_passes() == 2
    now[0] = 0.5
    assert fe.feed_audio(audio[16000:24000]) is False              # a pass is still running
    fe.decode_finished()
    assert fe.feed_audio(audio[24000:32000]) is False              # 0.5 s < decode interval
    now[0] = 1.0
    assert fe.feed_audio(audio[32000:40000]) is True               # interval reached
    fe.decode_finished()
    now[0] = 1.1
    decided = fe.feed_audio(audio[40000:140000])                   # crosses the 800-frame (8 s) window boundary
    assert fe.encoder.encoded_window_count == 1 and decided is True and fe.last_pass_is_boundary_finalize
    assert fe.min_agreement_passes() == 3                          # boundary boost active: stronger agreement
    fe.decode_finished()
    now[0] = 1.35
    assert fe.feed_audio(audio[140000:144000]) is True             # boundary cadence 0.2 s (last decode time was 1.0: finalize passes do not move it)
    fe.decode_finished()
    now[0] = 2.2
    assert fe.min_agreement_passes() == 2                          # boost (1 s) over
    assert fe.total_samples_fed == 144000


def test_no_finalize_mode_uses_the_interval_only(streaming):
    now = [0.0]
    mel = dsp.IncrementalMelSpectrogram(sample_rate=16000, n_fft=400, hop_length=160, n_mels=128)
    cfg = streaming.StreamingConfig(finalize_completed_windows=False, boundary_boost_seconds=0.0)
    fe = streaming.StreamingFrontEnd(mel, streaming.StreamingEncoder(FakeEncoder()), cfg, clock=lambda: now[0])
    audio = dsp.synth_audio(16000 * 9, 1)
    assert fe.feed_audio(audio[:8000]) is True
    fe.decode_finished()
    now[0] = 0.3
    assert fe.feed_audio(audio[8000:]) is False and fe.encoder.encoded_window_count == 1          # boundary alone does not force a pass
    now[0] = 1.0
    assert fe.feed_audio(np.zeros(1600, np.float32)) is True
