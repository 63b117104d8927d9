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


def test_library_sanitize_and_config_match_the_oracle(b2a, tmp_path):
    """csrc/weights.cu: b2a_weights_sanitize_speech_tokenizer / b2a_speech_tokenizer_config_from_json on a PyTorch-layout checkpoint
    written with the `safetensors` package."""
    import json
    from safetensors.torch import save_file
    cfg = oc.tiny_config()
    W = oc.init_weights(cfg, 9)
    ckpt = torch_layout_checkpoint(W)
    ckpt["speech_tokenizer.decoder.quantizer.rvq_first.vq.layers.0._codebook.initialized"] = torch.ones(1)
    ckpt["model.speaker_encoder.blocks.0.weight"] = torch.zeros(3, 3)
    d = tmp_path / "speech_tokenizer"
    d.mkdir()
    save_file({k: v.contiguous() for k, v in ckpt.items()}, str(d / "model.safetensors"))
    w = b2a.Weights(d)
    w.sanitize_speech_tokenizer()
    got = w.tensors()
    ref = oc.strip_decoder_prefix(oc.sanitize(ckpt))
    ref = {k: v for k, v in ref.items() if not k.endswith(".initialized")}
    assert set(got) == set(ref), set(got) ^ set(ref)
    for k, v in ref.items():
        assert got[k].shape == tuple(v.shape) and np.array_equal(got[k], v.numpy()), k
    # config: defaults when the file is missing, decoder_config overrides otherwise
    from mlx_audio_swift_b200.loading import speech_tokenizer_config_from_json
    c, rate = speech_tokenizer_config_from_json(tmp_path / "nope.json", max_batch=3, max_cache_frames=77)
    o = oc.TokenizerDecoderConfig()
    for f in ("codebook_size", "codebook_dim", "latent_dim", "decoder_dim", "hidden_size", "intermediate_size", "head_dim", "num_attention_heads",
              "num_key_value_heads", "num_hidden_layers", "num_quantizers", "num_semantic_quantizers"):
        assert getattr(c, f) == getattr(o, f), f
    assert abs(c.rms_norm_eps - 1e-5) < 1e-12 and c.rope_theta == 10000.0 and c.attention_bias == 0 and rate == 1920
    assert list(c.upsample_rates)[: c.num_upsample_rates] == [8, 5, 4, 3] and list(c.upsampling_ratios)[: c.num_upsampling_ratios] == [2, 2]
    assert (c.max_batch, c.max_cache_frames) == (3, 77)
    (d / "config.json").write_text(json.dumps({"decode_upsample_rate": 960, "decoder_config": {"latent_dim": 256, "upsample_rates": [4, 4], "attention_bias": True,
                                                                                                  "rope_theta": 5000.0, "sliding_window": 72}}))
    c, rate = speech_tokenizer_config_from_json(d / "config.json")
    assert c.latent_dim == 256 and c.codebook_dim == 512 and list(c.upsample_rates)[: c.num_upsample_rates] == [4, 4] and c.attention_bias == 1
    assert c.rope_theta == 5000.0 and rate == 960
    (d / "config.json").write_text(json.dumps({"decoder_config": {"upsample_rates": list(range(9))}}))
    with pytest.raises(b2a.AudioGenerationError) as e:
        speech_tokenizer_config_from_json(d / "config.json")
    assert e.value.case == "modelNotInitialized"
