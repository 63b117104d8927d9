"""Row N3: the streaming STT session's host logic (b2a_stt_session_*, whisper.cu) against oracle/stt_streaming.py
(StreamingInferenceSession.swift:589-950).  CPU tests drive it with a host-side decoder (the reference accepts `any
STTGenerationModel`); the GPU test runs it on the CUDA Whisper against the oracle session on the oracle Whisper."""
import numpy as np
import pytest

from oracle import stt_streaming as osx


def fake_decoder(stable: int):
    """Continuation after `prefix`: position i is fixed for i < stable + seconds heard, later positions flicker with the audio length."""
    def decode(audio, prefix):
        n = len(audio)
        k = min(60, n // 2500)
        out = []
        for i in range(len(prefix), k):
            settled = i < stable + n // 16000
            out.append((i * 7 + 3) % 31 if settled else (i * 5 + n // 4000) % 29)
        return out
    return decode


def run_both(b2a, cfg_kw, decoder, chunks, dt_scale=1.0):
    ocfg = osx.StreamingConfig(**cfg_kw)
    o = osx.StreamingSession(decoder, ocfg)
    c = b2a.StreamingInferenceSession(decoder=decoder, config=b2a.StreamingConfig(
        decode_interval_seconds=ocfg.decode_interval_s, window_seconds=ocfg.window_s, encoder_window_overlap_seconds=ocfg.window_overlap_s,
        delay_ms=ocfg.delay_ms, min_agreement_passes=ocfg.min_agreement_passes, max_tokens_per_pass=512, sample_rate=ocfg.sample_rate))
    now, kinds = 0.0, []
    for x in chunks:
        now += dt_scale * len(x) / 16000.0
        uo, uc = o.feed(x, now), c.feed_audio(x, now)
        assert uc.kind == uo.kind and uc.promoted == uo.promoted, (uc, uo)
        assert uc.completed == uo.completed and uc.confirmed == uo.confirmed and uc.provisional == uo.provisional
        assert abs(uc.total_audio_seconds - uo.total_audio_s) < 1e-9
        kinds.append(uc.kind)
    uo, uc = o.stop(now), c.stop(now)
    assert uc.kind == uo.kind == "ended"
    assert uc.completed == uo.completed and uc.confirmed == uo.confirmed and uc.provisional == uo.provisional == []
    assert c.feed_audio(np.zeros(100, np.float32), now + 1).kind == "none"          # a stopped session ignores audio
    return kinds, uc


@pytest.mark.parametrize("seed,cfg_kw", [(0, {}), (1, dict(delay_ms=200, min_agreement_passes=1)), (2, dict(decode_interval_s=0.3, delay_ms=2400)),
                                          (3, dict(window_s=3.0, window_overlap_s=0.5, min_agreement_passes=3)), (4, dict(window_overlap_s=9.0))])
def test_session_matches_the_oracle_on_random_feeds(b2a, seed, cfg_kw):
    rng = np.random.default_rng(seed)
    chunks = [rng.standard_normal(int(n)).astype(np.float32) for n in rng.integers(200, 12000, size=90)]
    kinds, last = run_both(b2a, cfg_kw, fake_decoder(stable=2), chunks)
    assert "partial" in kinds and "final_window" in kinds and len(last.completed) >= 2


def test_promotion_needs_delay_and_agreement(b2a):
    chunks = [np.zeros(8000, np.float32)] * 12                                       # 0.5 s per feed, one pass per second
    kinds, last = run_both(b2a, dict(delay_ms=480, min_agreement_passes=2), fake_decoder(stable=4), chunks)
    # the first pass can promote nothing (no agreement yet); later ones do
    o = osx.StreamingSession(fake_decoder(stable=4), osx.StreamingConfig())
    now, promoted = 0.0, []
    for x in chunks:
        now += 0.5
        u = o.feed(x, now)
        if u.kind == "partial":
            promoted.append(u.promoted)
    assert promoted[0] == 0 and sum(promoted) > 0


def test_empty_and_tiny_feeds_and_errors(b2a):
    s = b2a.StreamingInferenceSession(decoder=lambda a, p: [1, 2, 3])
    assert s.feed_audio(np.zeros(0, np.float32), 0.0).kind == "none"
    assert s.feed_audio(np.zeros(7999, np.float32), 0.1).kind == "none"               # < 0.5 s pending
    u = s.feed_audio(np.zeros(1, np.float32), 0.2)
    assert u.kind == "partial" and u.provisional == [1, 2, 3]
    assert s.stop(0.3).completed == [[1, 2, 3]]
    bad = b2a.StreamingInferenceSession(decoder=lambda a, p: 1 / 0)
    with pytest.raises(b2a.AudioGenerationError) as e:
        bad.feed_audio(np.zeros(9000, np.float32), 0.0)
    assert e.value.case == "generationFailed"
    with pytest.raises(b2a.AudioGenerationError):
        b2a.StreamingInferenceSession(decoder=lambda a, p: [], config=b2a.StreamingConfig(window_seconds=0.0))


@pytest.mark.gpu
def test_whisper_session_matches_the_oracle_session(b2a):
    import torch
    from oracle import dsp, whisper as ow
    cfg = ow.WhisperConfig(vocab_size=51865, num_mel_bins=80, d_model=64, encoder_layers=1, encoder_attention_heads=1, encoder_ffn_dim=128,
                           decoder_layers=1, decoder_attention_heads=1, decoder_ffn_dim=128)
    W = ow.init_weights(cfg, 5)
    hf = dict(vocab_size=cfg.vocab_size, num_mel_bins=80, d_model=64, encoder_layers=1, encoder_attention_heads=1, encoder_ffn_dim=128,
              max_source_positions=1500, decoder_layers=1, decoder_attention_heads=1, decoder_ffn_dim=128, max_target_positions=448)
    m = b2a.WhisperModel(hf, W, max_batch=1)
    P = b2a.STTGenerateParameters(max_tokens=6, mask_eot=True)
    prompt = ow.build_prompt_tokens()

    def oracle_decode(audio, prefix):
        o = ow.WhisperOracle(cfg, W)
        return ow.transcribe_tokens(o, np.asarray(audio, np.float32), prompt + list(prefix), max_tokens=6, mask_eot=True)

    scfg = dict(window_s=4.0, window_overlap_s=0.5, decode_interval_s=1.0, delay_ms=480, min_agreement_passes=2)
    o = osx.StreamingSession(oracle_decode, osx.StreamingConfig(**scfg))
    c = b2a.StreamingInferenceSession(m, b2a.StreamingConfig(decode_interval_seconds=1.0, window_seconds=4.0, encoder_window_overlap_seconds=0.5,
                                                            max_tokens_per_pass=6), P)
    x = dsp.synth_audio(16000 * 9, 31)
    now = 0.0
    seen = set()
    for i in range(0, len(x), 6000):
        now += 6000 / 16000.0
        uo, uc = o.feed(x[i:i + 6000], now), c.feed_audio(x[i:i + 6000], now)
        assert uc.kind == uo.kind and uc.completed == uo.completed and uc.confirmed == uo.confirmed and uc.provisional == uo.provisional, (i, uc, uo)
        seen.add(uc.kind)
    uo, uc = o.stop(now), c.stop(now)
    assert uc.completed == uo.completed and uc.confirmed == uo.confirmed
    assert {"partial", "final_window"} <= seen and len(uc.completed) >= 2
