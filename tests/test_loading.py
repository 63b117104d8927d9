"""Weight / format plumbing -- host-only code of the library, exercised through the C ABI on the CPU.
The checker is independent code: the `safetensors` package writes / reads the files, numpy restates the MLX affine
quantisation and the reference's key maps (WhisperModel.swift:315-480, LlamaTTS.swift:583-593) are spelled out as literal tables."""
import json

import numpy as np
import pytest
import torch
from safetensors.numpy import load_file, save_file
from safetensors.torch import save_file as save_file_torch


def test_safetensors_reader_matches_safetensors_package(b2a, tmp_path):
    rng = np.random.default_rng(0)
    a = {"f32": rng.standard_normal((3, 5)).astype(np.float32), "i32": rng.integers(-9, 9, (7,), dtype=np.int32),
         "f16": rng.standard_normal((2, 3, 4)).astype(np.float16), "i64": rng.integers(0, 100, (4,), dtype=np.int64)}
    save_file(a, str(tmp_path / "a.safetensors"))
    bf = torch.randn(6, 8).to(torch.bfloat16)
    save_file_torch({"bf16": bf, "f32": torch.ones(3, 5)}, str(tmp_path / "b.safetensors"))     # later file wins for "f32"
    w = b2a.Weights(tmp_path)
    t = w.tensors()
    assert set(t) == {"f32", "i32", "f16", "i64", "bf16"} and len(w) == 5
    assert np.array_equal(t["f32"], np.ones((3, 5), np.float32))                 # b.safetensors sorted after a.safetensors
    assert np.array_equal(t["i32"], a["i32"]) and np.array_equal(t["i64"], a["i64"].astype(np.int32))
    assert np.array_equal(t["f16"], a["f16"].astype(np.float32)) and t["f16"].dtype == np.float32
    assert torch.equal(t["bf16"], bf)
    single = b2a.Weights(tmp_path / "a.safetensors").tensors()
    ref = load_file(str(tmp_path / "a.safetensors"))
    assert np.array_equal(single["f32"], ref["f32"])
    with pytest.raises(b2a.AudioGenerationError) as e:
        b2a.Weights(tmp_path / "missing")
    assert e.value.case == "modelNotInitialized"


HF_KEYS = ["model.encoder.conv1.weight", "model.encoder.conv1.bias", "model.encoder.conv2.weight", "model.encoder.embed_positions.weight",
           "model.encoder.layers.0.self_attn.q_proj.weight", "model.encoder.layers.0.self_attn_layer_norm.bias",
           "model.encoder.layers.0.fc1.weight", "model.encoder.layer_norm.weight", "model.decoder.embed_tokens.weight",
           "model.decoder.embed_positions.weight", "model.decoder.layers.1.encoder_attn.out_proj.bias",
           "model.decoder.layers.1.encoder_attn_layer_norm.weight", "model.decoder.layers.1.final_layer_norm.weight",
           "model.decoder.layers.1.fc2.bias", "model.decoder.layer_norm.bias"]
MLX_KEYS = ["encoder.conv1.weight", "encoder.conv1.bias", "encoder.conv2.weight", None,
            "encoder.blocks.0.attn.query.weight", "encoder.blocks.0.attn_ln.bias", "encoder.blocks.0.mlp1.weight", "encoder.ln_post.weight",
            "decoder.token_embedding.weight", "decoder.positional_embedding", "decoder.blocks.1.cross_attn.out.bias",
            "decoder.blocks.1.cross_attn_ln.weight", "decoder.blocks.1.mlp_ln.weight", "decoder.blocks.1.mlp2.bias", "decoder.ln.bias"]


def whisper_values(rng, d=8):
    shapes = {"conv1.weight": (d, 3, 5), "conv2.weight": (d, 3, d), "embed_positions.weight": (1500, d)}
    vals = {}
    for k in HF_KEYS:
        shp = next((s for suf, s in shapes.items() if k.endswith(suf) and "encoder" in k), (d,))
        vals[k] = rng.standard_normal(shp).astype(np.float32)
    return vals


def test_whisper_sanitize_hugging_face(b2a, tmp_path):
    vals = whisper_values(np.random.default_rng(1))
    raw = {(k[len("model."):] if i % 2 else k): v for i, (k, v) in enumerate(vals.items())}      # re-exports drop "model."
    raw["proj_out.weight"] = np.zeros((4, 8), np.float32)                                        # tied: dropped
    save_file(raw, str(tmp_path / "model.safetensors"))
    w = b2a.Weights(tmp_path)
    assert w.sanitize_whisper() == 0
    t = w.tensors()
    assert set(t) == set(HF_KEYS)
    for k in HF_KEYS:
        assert np.array_equal(t[k], vals[k]), k          # conv weights stay in the PyTorch [out, in, k] layout b2a_stt_create takes


def test_whisper_sanitize_mlx_whisper_and_sinusoids(b2a, tmp_path):
    rng = np.random.default_rng(2)
    vals = whisper_values(rng)
    raw = {}
    for hf, mk in zip(HF_KEYS, MLX_KEYS):
        if mk is None:
            continue                                          # mlx-whisper omits the encoder positions
        v = vals[hf]
        if hf.endswith("conv1.weight") or hf.endswith("conv2.weight"):
            v = np.ascontiguousarray(v.transpose(0, 2, 1))    # stored in MLX's [out, k, in]
        raw[mk] = v
    raw["alignment_heads"] = np.zeros((2, 2), np.int32)
    raw["decoder.blocks.1.unknown.weight"] = np.zeros(3, np.float32)       # unmapped keys are dropped (remapMlxWhisperKey -> nil)
    save_file(raw, str(tmp_path / "weights.safetensors"))
    w = b2a.Weights(tmp_path)
    assert w.sanitize_whisper() == 1
    t = w.tensors()
    assert set(t) == set(HF_KEYS)
    for k in HF_KEYS:
        if k != "model.encoder.embed_positions.weight":
            assert np.array_equal(t[k], vals[k]), k
    # whisperSinusoids(length: 1500, channels: conv2.shape[0])  (WhisperModel.swift:381-395)
    ch, half = 8, 4
    inc = np.log(10000.0) / (half - 1)
    st = np.arange(1500)[:, None] * np.exp(-inc * np.arange(half))[None, :]
    ref = np.concatenate([np.sin(st), np.cos(st)], axis=1).astype(np.float32)
    assert t["model.encoder.embed_positions.weight"].shape == (1500, ch)
    assert np.abs(t["model.encoder.embed_positions.weight"] - ref).max() < 1e-6


def mlx_affine_quantize(w, group_size, bits):
    """numpy restatement of MLX's affine group quantisation (mx.quantize): per group of `group_size` input columns,
    scale = (max - min) / (2^bits - 1), bias = min, q = round((w - bias) / scale); 32/bits values per uint32, low bits first."""
    rows, cols = w.shape
    g = w.reshape(rows, cols // group_size, group_size).astype(np.float64)
    lo, hi = g.min(-1, keepdims=True), g.max(-1, keepdims=True)
    scale = np.where(hi > lo, (hi - lo) / (2 ** bits - 1), 1.0)
    q = np.clip(np.rint((g - lo) / scale), 0, 2 ** bits - 1).astype(np.uint32).reshape(rows, cols)
    per = 32 // bits
    words = np.zeros((rows, cols // per), dtype=np.uint32)
    for j in range(per):
        words |= q[:, j::per] << np.uint32(j * bits)
    return words, scale[..., 0].astype(np.float32), lo[..., 0].astype(np.float32), q


@pytest.mark.parametrize("bits,group_size", [(4, 64), (8, 32), (2, 64)])
def test_llama_sanitize_and_mlx_affine_dequant(b2a, tmp_path, bits, group_size):
    rng = np.random.default_rng(bits)
    wq = rng.standard_normal((16, 128)).astype(np.float32)
    words, scales, biases, q = mlx_affine_quantize(wq, group_size, bits)
    plain = torch.randn(8, 16).to(torch.bfloat16)
    save_file({"model.layers.0.mlp.down_proj.weight": words.view(np.int32), "model.layers.0.mlp.down_proj.scales": scales,
               "model.layers.0.mlp.down_proj.biases": biases, "model.layers.0.self_attn.rotary_emb.inv_freq": np.ones(4, np.float32),
               "lm_head.weight": np.ones((2, 2), np.float32)}, str(tmp_path / "model-00001.safetensors"))
    save_file_torch({"model.norm.weight": plain}, str(tmp_path / "model-00002.safetensors"))
    w = b2a.Weights(tmp_path)
    w.sanitize_llama(True, group_size, bits)
    t = w.tensors()
    assert set(t) == {"model.layers.0.mlp.down_proj.weight", "model.norm.weight"}        # inv_freq, tied lm_head, scales, biases gone
    deq = t["model.layers.0.mlp.down_proj.weight"]
    assert deq.dtype == torch.bfloat16 and tuple(deq.shape) == (16, 128)
    ref = (np.repeat(scales, group_size, axis=1) * q + np.repeat(biases, group_size, axis=1)).astype(np.float32)
    assert torch.equal(deq, torch.from_numpy(ref).to(torch.bfloat16))                     # w = scales * q + biases, rounded to bf16
    assert torch.equal(t["model.norm.weight"], plain)
    w2 = b2a.Weights(tmp_path)
    w2.sanitize_llama(False, 0, 0)
    assert "lm_head.weight" in w2.tensors() and "model.layers.0.mlp.down_proj.scales" in w2.tensors()
    with pytest.raises(b2a.AudioGenerationError):
        b2a.Weights(tmp_path).sanitize_llama(True, 64, 3)


def test_quantized_whisper_checkpoint_is_dequantized_after_the_remap(b2a, tmp_path):
    """WhisperModel.fromDirectory quantises every Linear and decoder.embed_tokens when config.json has "quantization"
    (WhisperModel.swift:499-511); Tests/WhisperQuantizedTiedEmbeddingTests.swift pins that the tied projection then uses the
    DE-QUANTISED embedding (tolerance 1e-2).  Here: an mlx-whisper-layout checkpoint with a packed token embedding and a packed Linear."""
    rng = np.random.default_rng(5)
    emb = rng.standard_normal((96, 64)).astype(np.float32)
    fc1 = rng.standard_normal((128, 64)).astype(np.float32)
    we, se, be, qe = mlx_affine_quantize(emb, 32, 4)
    wf, sf, bf_, qf = mlx_affine_quantize(fc1, 32, 4)
    save_file({"decoder.token_embedding.weight": we.view(np.int32), "decoder.token_embedding.scales": se, "decoder.token_embedding.biases": be,
               "decoder.blocks.0.mlp1.weight": wf.view(np.int32), "decoder.blocks.0.mlp1.scales": sf, "decoder.blocks.0.mlp1.biases": bf_,
               "decoder.blocks.0.mlp1.bias": np.zeros(128, np.float32), "encoder.conv2.weight": rng.standard_normal((64, 3, 64)).astype(np.float32)},
              str(tmp_path / "weights.safetensors"))
    w = b2a.Weights(tmp_path)
    assert w.sanitize_whisper() == 1                                # mlx-whisper layout detected
    w.dequantize(32, 4)
    t = w.tensors()
    assert "model.decoder.embed_tokens.scales" not in t and "model.decoder.layers.0.fc1.biases" not in t
    assert "model.decoder.layers.0.fc1.bias" in t                  # the layer's own bias is not the quantiser's "biases"
    dense_e = (np.repeat(se, 32, axis=1) * qe + np.repeat(be, 32, axis=1)).astype(np.float32)
    dense_f = (np.repeat(sf, 32, axis=1) * qf + np.repeat(bf_, 32, axis=1)).astype(np.float32)
    assert torch.equal(t["model.decoder.embed_tokens.weight"], torch.from_numpy(dense_e).to(torch.bfloat16))
    assert torch.equal(t["model.decoder.layers.0.fc1.weight"], torch.from_numpy(dense_f).to(torch.bfloat16))
    hidden = rng.standard_normal(64).astype(np.float32)            # the reference test's statement, with its tolerance
    assert np.abs(t["model.decoder.embed_tokens.weight"].float().numpy() @ hidden - dense_e @ hidden).max() < 1e-2 * max(1.0, np.abs(dense_e @ hidden).max())
    with pytest.raises(b2a.AudioGenerationError):
        b2a.Weights(tmp_path).dequantize(32, 0)


def test_llama_per_layer_quantization_from_config(b2a, tmp_path):
    """config.json "quantization" with per-layer overrides (mlx-swift-lm PerLayerQuantization, call site LlamaTTS.swift:955-966):
    the default 4-bit / 64, one layer at 8-bit / 32, one layer marked false."""
    rng = np.random.default_rng(11)
    a, b, c = (rng.standard_normal((16, 128)).astype(np.float32) for _ in range(3))
    wa, sa, ba, qa = mlx_affine_quantize(a, 64, 4)
    wb, sb, bb, qb = mlx_affine_quantize(b, 32, 8)
    tensors = {"model.layers.0.mlp.down_proj.weight": wa.view(np.int32), "model.layers.0.mlp.down_proj.scales": sa, "model.layers.0.mlp.down_proj.biases": ba,
               "model.embed_tokens.weight": wb.view(np.int32), "model.embed_tokens.scales": sb, "model.embed_tokens.biases": bb,
               "model.layers.0.mlp.up_proj.weight": c, "lm_head.weight": np.ones((2, 2), np.float32)}
    save_file(tensors, str(tmp_path / "model.safetensors"))
    cfg = {"tie_word_embeddings": True, "quantization": {"group_size": 64, "bits": 4, "model.embed_tokens": {"group_size": 32, "bits": 8},
                                                          "model.layers.0.mlp.up_proj": False}}
    (tmp_path / "config.json").write_text(json.dumps(cfg))
    w = b2a.Weights(tmp_path)
    w.sanitize_llama_config(tmp_path / "config.json")
    t = w.tensors()
    assert set(t) == {"model.layers.0.mlp.down_proj.weight", "model.embed_tokens.weight", "model.layers.0.mlp.up_proj.weight"}
    ref_a = (np.repeat(sa, 64, axis=1) * qa + np.repeat(ba, 64, axis=1)).astype(np.float32)
    ref_b = (np.repeat(sb, 32, axis=1) * qb + np.repeat(bb, 32, axis=1)).astype(np.float32)
    assert torch.equal(t["model.layers.0.mlp.down_proj.weight"], torch.from_numpy(ref_a).to(torch.bfloat16))
    assert torch.equal(t["model.embed_tokens.weight"], torch.from_numpy(ref_b).to(torch.bfloat16))
    assert np.array_equal(t["model.layers.0.mlp.up_proj.weight"], c)                      # untouched fp32
    # a layer the config marks unquantised must not carry scales
    cfg["quantization"]["model.layers.0.mlp.down_proj"] = False
    (tmp_path / "config.json").write_text(json.dumps(cfg))
    with pytest.raises(b2a.AudioGenerationError) as e:
        b2a.Weights(tmp_path).sanitize_llama_config(tmp_path / "config.json")
    assert e.value.case == "modelNotInitialized"
    # no "quantization" at all: tensors pass through (the tied lm_head is still dropped)
    (tmp_path / "config.json").write_text(json.dumps({"tie_word_embeddings": True}))
    w3 = b2a.Weights(tmp_path)
    w3.sanitize_llama_config(tmp_path / "config.json")
    assert "model.embed_tokens.scales" in w3.tensors() and "lm_head.weight" not in w3.tensors()


def test_llama_config_from_json(b2a, tmp_path):
    cfg = {"hidden_size": 3072, "num_hidden_layers": 28, "intermediate_size": 8192, "num_attention_heads": 24, "num_key_value_heads": 8,
           "rms_norm_eps": 1e-5, "vocab_size": 156940, "rope_theta": 500000.0, "tie_word_embeddings": True, "model_type": "llama",
           "rope_scaling": {"factor": 32.0, "high_freq_factor": 4.0, "low_freq_factor": 1.0, "original_max_position_embeddings": 8192,
                            "rope_type": "llama3"}, "quantization": {"group_size": 64, "bits": 4}, "torch_dtype": "bfloat16", "eos_token_id": [1, 2]}
    p = tmp_path / "config.json"
    p.write_text(json.dumps(cfg))
    c, gs, bits = b2a.llama_config_from_json(p, max_batch=8, max_context=640)
    assert (c.hidden_size, c.num_hidden_layers, c.intermediate_size, c.num_attention_heads, c.num_key_value_heads, c.head_dim,
            c.vocab_size, c.tie_word_embeddings, c.max_batch, c.max_context) == (3072, 28, 8192, 24, 8, 128, 156940, 1, 8, 640)
    assert abs(c.rope_theta - 500000.0) < 1 and (c.rope_factor, c.rope_low_freq_factor, c.rope_high_freq_factor, c.rope_old_context_len) == (32.0, 1.0, 4.0, 8192.0)
    assert (gs, bits) == (64, 4)
    bad = dict(cfg); bad["rope_scaling"] = {"rope_type": "llama3"}              # rope_scaling must contain 'factor' (LlamaTTSConfig.swift:139-144)
    p.write_text(json.dumps(bad))
    with pytest.raises(b2a.AudioGenerationError):
        b2a.llama_config_from_json(p)
    del cfg["vocab_size"]
    p.write_text(json.dumps(cfg))
    with pytest.raises(b2a.AudioGenerationError):
        b2a.llama_config_from_json(p)


def test_qwen3_talker_config_and_sanitize(b2a, tmp_path):
    """Qwen3TTSModel.fromModelDirectory's talker half, host side: talker_config / code_predictor_config decoding with the
    reference's defaults (Qwen3TTSConfig.swift:45-63,268-292), sanitize keeps "talker.*" without the prefix
    (Qwen3TTSTalker.swift:356-365), 8-bit layers are expanded (Qwen3TTS.swift:1156-1171)."""
    import ctypes as C
    f = b2a._ffi
    (tmp_path / "config.json").write_text(json.dumps({"model_type": "qwen3_tts", "quantization": {"group_size": 64, "bits": 8},
                                                      "talker_config": {"hidden_size": 512, "num_hidden_layers": 3, "codec_eos_token_id": 2151,
                                                                        "code_predictor_config": {"num_hidden_layers": 2, "vocab_size": 1024}}}))
    c = f.Qwen3TalkerConfig()
    f.check(f.lib().b2a_qwen3_talker_config_from_json(str(tmp_path / "config.json").encode(), 4, 300, C.byref(c)))
    assert (c.hidden_size, c.num_hidden_layers, c.vocab_size, c.codec_eos_token_id, c.text_vocab_size) == (512, 3, 3072, 2151, 151936)
    assert (c.cp_num_hidden_layers, c.cp_vocab_size, c.cp_hidden_size, c.num_code_groups, c.max_batch, c.max_context) == (2, 1024, 1024, 16, 4, 300)
    assert abs(c.rope_theta - 1e6) < 1 and abs(c.rms_norm_eps - 1e-6) < 1e-12 and c.head_dim == 128
    rng = np.random.default_rng(8)
    wq = rng.standard_normal((16, 128)).astype(np.float32)
    words, scales, biases, q = mlx_affine_quantize(wq, 64, 8)
    save_file({"talker.model.layers.0.mlp.down_proj.weight": words.view(np.int32), "talker.model.layers.0.mlp.down_proj.scales": scales,
               "talker.model.layers.0.mlp.down_proj.biases": biases, "talker.model.norm.weight": np.ones(4, np.float32),
               "speaker_encoder.fc.weight": np.ones((2, 2), np.float32)}, str(tmp_path / "model.safetensors"))
    w = b2a.Weights(tmp_path)
    f.check(f.lib().b2a_weights_sanitize_qwen3_talker(w._h, str(tmp_path / "config.json").encode()))
    t = w.tensors()
    assert set(t) == {"model.layers.0.mlp.down_proj.weight", "model.norm.weight"}
    ref = (np.repeat(scales, 64, axis=1) * q + np.repeat(biases, 64, axis=1)).astype(np.float32)
    assert torch.equal(t["model.layers.0.mlp.down_proj.weight"], torch.from_numpy(ref).to(torch.bfloat16))
