"""Host side of the (experimental) Qwen3-TTS speech-tokenizer mirror: sanitize and config handling against the oracle's
restatement, and the no-CPU-fallback rule.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import qwen3_tts_codec as oc


@pytest.fixture(scope="module")
def codec():
    import importlib
    return importlib.import_module("mlx_audio_swift_b200.qwen3_tts_codec")


def torch_layout_checkpoint(W):
    ckpt = {}
    for k, v in W.items():
        key = "decoder." + k
        if ".codebook." in key:
            key = key.replace(".codebook.", "._codebook.")
        if k.startswith("upsample."):
            key = key.replace(".layers.", ".")
        is_t = (k.startswith("upsample.") and k.endswith("layers.0.conv.weight")) or (k.startswith("decoder.") and k.endswith("block.1.conv.weight"))
        if v.ndim == 3:
            v = v.permute(2, 0, 1) if is_t else v.permute(0, 2, 1)
        ckpt["speech_tokenizer." + key] = v.contiguous()
    ckpt["speech_tokenizer.encoder.encoder.layers.0.conv.weight"] = torch.zeros(4, 1, 7)
    ckpt["speaker_encoder.fc.weight"] = torch.zeros(2, 2)
    return ckpt


def test_sanitize_matches_the_oracle_restatement(codec):
    cfg = oc.tiny_config()
    ckpt = torch_layout_checkpoint(oc.init_weights(cfg, 9))
    ref = oc.sanitize(ckpt)
    got = codec.sanitize({k: v.numpy() for k, v in ckpt.items()})
    ref = {k: v for k, v in ref.items() if not k.endswith(".initialized")}
    assert set(got) == set(ref)
    for k in ref:
        assert got[k].shape == tuple(ref[k].shape) and np.array_equal(got[k], ref[k].numpy()), k
    assert codec.check_array_shape((1536, 7, 1024)) and not codec.check_array_shape((1536, 1024, 7))


def test_config_defaults_and_from_dict(codec):
    c = codec.Qwen3TTSTokenizerDecoderConfig()
    o = oc.TokenizerDecoderConfig()
    for f in c.__dataclass_fields__:
        assert getattr(c, f) == getattr(o, f), f
    c2 = codec.Qwen3TTSTokenizerDecoderConfig.from_dict({"latent_dim": 64, "sliding_window": 72, "hidden_act": "silu"})
    assert c2.latent_dim == 64 and c2.codebook_dim == 512


def test_create_fails_loudly_without_a_device(b2a, codec):
    if b2a.device_count() > 0:
        pytest.skip("needs a box without a GPU")
    cfg = oc.tiny_config(decoder_dim=128, head_dim=32, num_attention_heads=2, num_key_value_heads=1)
    W = {k: v.numpy() for k, v in oc.init_weights(cfg, 1).items()}
    with pytest.raises(b2a.AudioGenerationError) as e:
        codec.Qwen3TTSSpeechTokenizerDecoder(codec.Qwen3TTSTokenizerDecoderConfig.from_dict({k: getattr(cfg, k) for k in cfg.__dataclass_fields__}), weights=W)
    assert e.value.case == "cudaError"
    bad = oc.tiny_config()                                        # decoder_dim / 2^4 = 4 channels: rejected before any device work
    with pytest.raises(b2a.AudioGenerationError) as e:
        codec.Qwen3TTSSpeechTokenizerDecoder(codec.Qwen3TTSTokenizerDecoderConfig.from_dict({k: getattr(bad, k) for k in bad.__dataclass_fields__}), weights=W)
    assert e.value.case == "invalidInput"
