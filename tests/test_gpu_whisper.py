"""CUDA Whisper (through the C ABI) vs the oracle: encoder hidden states and decoder logits within 1e-3 relative
(fp32-activation oracle on the same bf16 weights), greedy token ids bit-exact, suppression / stop semantics,
batched == serial."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err
from oracle import dsp
from oracle import whisper as ow

pytestmark = pytest.mark.gpu
TOL = 1e-3


def hf_config(cfg: ow.WhisperConfig) -> dict:
    return dict(vocab_size=cfg.vocab_size, num_mel_bins=cfg.num_mel_bins, d_model=cfg.d_model, encoder_layers=cfg.encoder_layers,
                encoder_attention_heads=cfg.encoder_attention_heads, encoder_ffn_dim=cfg.encoder_ffn_dim,
                max_source_positions=1500, decoder_layers=cfg.decoder_layers,
                decoder_attention_heads=cfg.decoder_attention_heads, decoder_ffn_dim=cfg.decoder_ffn_dim,
                max_target_positions=448)


@pytest.fixture(scope="module")
def tiny(b2a):
    cfg = ow.WhisperConfig.tiny_test()
    W = ow.init_weights(cfg, 1234)
    return cfg, W, b2a.WhisperModel(hf_config(cfg), W, max_batch=4)


def test_encoder_vs_oracle_and_golden(tiny):
    cfg, W, m = tiny
    xs = np.stack([np.pad(dsp.synth_audio(64000, 3), (0, 16000)), dsp.synth_audio(80000, 4)])
    enc = m.encode(xs)
    assert enc.shape == (2, 1500, cfg.d_model)
    o = ow.WhisperOracle(cfg, W)
    for i in range(2):
        ref = o.encode(torch.from_numpy(dsp.whisper_encoder_features(xs[i])).float()).numpy()[0]
        assert rel_err(enc[i], ref) < TOL, rel_err(enc[i], ref)
    g = np.load(GOLDEN / "whisper_tiny.npz")
    one = m.encode(dsp.synth_audio(64000, 3))
    assert rel_err(one[0, [0, 1, 700, 1499]], g["enc_rows"]) < TOL
    assert rel_err(one[0], enc[0]) < 1e-5                       # zero-padding to 30 s is what padOrTrimToWindow does


def test_decoder_logits_vs_oracle(tiny):
    cfg, W, m = tiny
    x = dsp.synth_audio(64000, 3)
    m.encode(x)
    ids = np.asarray([ow.build_prompt_tokens() + [11, 2222, 33333, 4, 50000]], dtype=np.int32)
    lg = m.decoder_logits(ids)
    o = ow.WhisperOracle(cfg, W)
    enc = o.encode(torch.from_numpy(dsp.whisper_encoder_features(x)).float())
    ref = o.logits(o.decode(torch.as_tensor(ids, dtype=torch.long), 0, enc)).numpy()
    assert lg.shape == ref.shape and rel_err(lg, ref) < TOL, rel_err(lg, ref)
    assert np.array_equal(lg.argmax(-1), ref.argmax(-1))


def test_greedy_tokens_bit_exact_and_batched_equals_serial(b2a, tiny):
    cfg, W, m = tiny
    g = np.load(GOLDEN / "whisper_tiny.npz")
    P = b2a.STTGenerateParameters(max_tokens=12, mask_eot=True)
    x3 = dsp.synth_audio(64000, 3)
    out = m.generate(x3, P)
    assert out.tokens[0] == g["greedy"].tolist()
    assert out.prompt_tokens == 4 and out.generation_tokens == 12 and out.total_time > 0
    xs = np.stack([np.pad(x3, (0, 16000)), dsp.synth_audio(80000, 4), dsp.synth_audio(80000, 5)])
    outb = m.generate(xs, P)
    assert outb.tokens[0] == out.tokens[0]
    for i in (1, 2):
        ref = ow.transcribe_tokens(ow.WhisperOracle(cfg, W), xs[i], ow.build_prompt_tokens(), max_tokens=12, mask_eot=True)
        assert outb.tokens[i] == ref


def test_natural_stop_and_suppression(b2a, tiny):
    cfg, W, m = tiny
    x = dsp.synth_audio(48000, 9)
    P = b2a.STTGenerateParameters(max_tokens=40)
    out = m.generate(x, P)
    ref = ow.transcribe_tokens(ow.WhisperOracle(cfg, W), x, ow.build_prompt_tokens(), max_tokens=40)
    assert out.tokens[0] == ref and all(t < ow.TIMESTAMP_BEGIN and t != ow.EOT for t in out.tokens[0])
    # a suppress list changes the pick exactly like the oracle's additive -1e9 mask
    banned = ref[:3]
    P2 = b2a.STTGenerateParameters(max_tokens=10, suppress_tokens=banned, mask_eot=True)
    out2 = m.generate(x, P2)
    ref2 = ow.transcribe_tokens(ow.WhisperOracle(cfg, W), x, ow.build_prompt_tokens(), max_tokens=10, suppress=banned, mask_eot=True)
    assert out2.tokens[0] == ref2 and not set(out2.tokens[0]) & set(banned)
    # translate task / no language: different prefix lengths
    P3 = b2a.STTGenerateParameters(max_tokens=6, language_id=None, task="translate", mask_eot=True)
    out3 = m.generate(x, P3)
    ref3 = ow.transcribe_tokens(ow.WhisperOracle(cfg, W), x, ow.build_prompt_tokens(None, "translate"), max_tokens=6, mask_eot=True)
    assert out3.tokens[0] == ref3 and out3.prompt_tokens == 3


def test_temperature_sampling_is_categorical_over_the_masked_logits(b2a, tiny):
    """WhisperModel.swift:284-291: temperature > 0 -> categorical(logits / T) after the suppress masks.  The draw stream cannot match
    MLX's generator; the DISTRIBUTION is checked against the oracle's masked first-step logits, plus determinism per seed."""
    cfg, W, m = tiny
    x = dsp.synth_audio(48000, 9)
    o = ow.WhisperOracle(cfg, W)
    enc = o.encode(torch.from_numpy(dsp.whisper_encoder_features(x)).float())
    prompt = ow.build_prompt_tokens()
    lg = o.logits(o.decode(torch.as_tensor([prompt], dtype=torch.long), 0, enc)).numpy()[0, -1].astype(np.float64)
    lg[ow.EOT] += -1e9                                  # begin-suppress (step 0), timestamps always
    lg[ow.TIMESTAMP_BEGIN:] += -1e9
    srt = np.sort(lg)
    T = float(max(1e-3, (srt[-1] - srt[-8]) / 1.5))     # scaled to the logits: the top handful of tokens share most of the mass
    p = np.exp((lg - lg.max()) / T); p /= p.sum()
    draws = [m.generate(x, b2a.STTGenerateParameters(max_tokens=1, temperature=T, seed=s)).tokens[0][0] for s in range(300)]
    assert all(t < ow.TIMESTAMP_BEGIN and t != ow.EOT for t in draws)
    assert len(set(draws)) > 3
    top = int(np.argmax(p))
    f, e = draws.count(top) / len(draws), p[top]
    assert abs(f - e) < 4 * np.sqrt(e * (1 - e) / len(draws)) + 0.02, (f, e)
    heavy = np.argsort(-p)[:20]
    assert sum(d in set(heavy.tolist()) for d in draws) / len(draws) > p[heavy].sum() - 0.08
    a = m.generate(x, b2a.STTGenerateParameters(max_tokens=6, temperature=T, seed=5, mask_eot=True)).tokens
    b = m.generate(x, b2a.STTGenerateParameters(max_tokens=6, temperature=T, seed=5, mask_eot=True)).tokens
    c = m.generate(x, b2a.STTGenerateParameters(max_tokens=6, temperature=T, seed=6, mask_eot=True)).tokens
    assert a == b and a != c
    cold = m.generate(x, b2a.STTGenerateParameters(max_tokens=6, temperature=1e-4, seed=3, mask_eot=True)).tokens
    assert cold == m.generate(x, b2a.STTGenerateParameters(max_tokens=6, mask_eot=True)).tokens      # T -> 0 is the argmax


def test_errors(b2a, tiny):
    cfg, W, m = tiny
    with pytest.raises(b2a.AudioGenerationError) as e:
        m.generate(np.zeros((5, 16000), np.float32))           # max_batch 4
    assert e.value.case == "invalidInput"
    bad = dict(hf_config(cfg)); bad["encoder_attention_heads"] = 4
    with pytest.raises(b2a.AudioGenerationError):
        b2a.WhisperModel(bad, W)


def test_long_audio_is_chunked_into_30s_windows_and_batched(b2a, tiny):
    # WhisperModel.swift:165-182: consecutive 30 s windows, last one shorter; each window == transcribing that slice alone
    cfg, W, m = tiny
    x = np.concatenate([dsp.synth_audio(480000, 5), dsp.synth_audio(480000, 6), dsp.synth_audio(160000, 7)])      # 70 s
    P = b2a.STTGenerateParameters(max_tokens=6, mask_eot=True)
    out = m.generate_long(x, P)
    assert len(out.tokens) == 3 and [s["start"] for s in out.segments] == [0.0, 30.0, 60.0] and out.segments[2]["end"] == 70.0
    for i, (lo, hi) in enumerate(((0, 480000), (480000, 960000), (960000, 1120000))):
        assert out.tokens[i] == m.generate(x[lo:hi][None], P).tokens[0]
    short = m.generate_long(x[:1000], P)                                   # <= one window: a single chunk (:169-171)
    assert len(short.tokens) == 1 and short.segments[0]["start"] == 0.0
    stereo = np.stack([x[:200000], x[:200000]], axis=-1)
    assert m.generate_long(stereo, P).tokens == m.generate_long(x[:200000], P).tokens       # mono = mean over channels (:98)
