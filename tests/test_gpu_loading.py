"""from_model_directory through the library's loader on the device: a Whisper checkpoint in the HF and in the
mlx-whisper key layout gives the same encoder states as the dict-constructed model; an Orpheus-style directory (sharded
safetensors, one 4-bit MLX-affine-quantised projection) gives the logits of the model built from the de-quantised weights."""
import json

import numpy as np
import pytest
import torch
from safetensors.numpy import save_file
from safetensors.torch import save_file as save_file_torch

from oracle import dsp
from oracle import llama as ol
from oracle import whisper as ow
from test_gpu_llama import TINY, hf_config as llama_hf_config
from test_gpu_whisper import hf_config as whisper_hf_config
from test_loading import mlx_affine_quantize

pytestmark = pytest.mark.gpu

HF_TO_MLX = [("model.encoder.embed_positions.weight", None), ("model.decoder.embed_positions.weight", "decoder.positional_embedding"),
             ("model.decoder.embed_tokens.", "decoder.token_embedding."), ("model.encoder.layer_norm.", "encoder.ln_post."),
             ("model.decoder.layer_norm.", "decoder.ln."), ("model.encoder.conv", "encoder.conv")]
SUFFIX = [("self_attn_layer_norm.", "attn_ln."), ("encoder_attn_layer_norm.", "cross_attn_ln."), ("final_layer_norm.", "mlp_ln."),
          ("fc1.", "mlp1."), ("fc2.", "mlp2."), ("self_attn.q_proj.", "attn.query."), ("self_attn.k_proj.", "attn.key."),
          ("self_attn.v_proj.", "attn.value."), ("self_attn.out_proj.", "attn.out."), ("encoder_attn.q_proj.", "cross_attn.query."),
          ("encoder_attn.k_proj.", "cross_attn.key."), ("encoder_attn.v_proj.", "cross_attn.value."), ("encoder_attn.out_proj.", "cross_attn.out.")]


def to_mlx_whisper(W):
    out = {}
    for k, v in W.items():
        v = np.asarray(v, dtype=np.float32)
        for stem in ("encoder", "decoder"):
            pre = f"model.{stem}.layers."
            if k.startswith(pre):
                idx, rest = k[len(pre):].split(".", 1)
                for a, b in SUFFIX:
                    if rest.startswith(a):
                        out[f"{stem}.blocks.{idx}.{b}{rest[len(a):]}"] = v
                        break
                break
        else:
            for a, b in HF_TO_MLX:
                if k.startswith(a):
                    if b is not None:
                        out[b + k[len(a):]] = np.ascontiguousarray(v.transpose(0, 2, 1)) if k.endswith("conv1.weight") or k.endswith("conv2.weight") else v
                    break
    return out


def test_whisper_from_model_directory_both_formats(b2a, tmp_path):
    cfg = ow.WhisperConfig.tiny_test()
    W = {k: v.to(torch.float32).numpy() for k, v in ow.init_weights(cfg, 1234).items()}      # bf16-valued matrices widened
    base = b2a.WhisperModel(whisper_hf_config(cfg), W, max_batch=2)
    x = dsp.synth_audio(64000, 3)
    ref = base.encode(x)
    hf_dir, mlx_dir = tmp_path / "hf", tmp_path / "mlx"
    hf_dir.mkdir(); mlx_dir.mkdir()
    (hf_dir / "config.json").write_text(json.dumps(whisper_hf_config(cfg)))
    save_file(W, str(hf_dir / "model.safetensors"))
    (mlx_dir / "config.json").write_text(json.dumps(whisper_hf_config(cfg)))
    save_file(to_mlx_whisper(W), str(mlx_dir / "weights.safetensors"))
    a = b2a.WhisperModel.from_model_directory(hf_dir, max_batch=2).encode(x)
    assert np.array_equal(a, ref)
    m2 = b2a.WhisperModel.from_model_directory(mlx_dir, max_batch=2)
    b = m2.encode(x)
    # mlx-whisper omits the encoder positions: the loader synthesises the sinusoid the oracle's checkpoint holds
    assert np.abs(W["model.encoder.embed_positions.weight"] - ow.sinusoids(1500, cfg.d_model).numpy()).max() < 1e-6
    assert np.abs(a - b).max() < 1e-5 * max(1.0, np.abs(a).max())


def test_llama_from_model_directory_sharded_and_quantised(b2a, tmp_path):
    cfg = ol.LlamaConfig(**TINY)
    W = ol.init_weights(cfg, 7, std=0.08)
    W = {k: (v if isinstance(v, torch.Tensor) else torch.as_tensor(v)) for k, v in W.items()}
    qname = "model.layers.1.mlp.down_proj"
    w32 = W[qname + ".weight"].to(torch.float32).numpy()
    words, scales, biases, q = mlx_affine_quantize(w32, 64, 4)
    deq = (np.repeat(scales, 64, axis=1) * q + np.repeat(biases, 64, axis=1)).astype(np.float32)
    Wd = dict(W); Wd[qname + ".weight"] = torch.from_numpy(deq).to(torch.bfloat16)
    ref_model = b2a.LlamaTTSModel(llama_hf_config(cfg), Wd, max_batch=2, max_context=64)
    ids = np.asarray([[5, 17, 99, 4, 1000, 3], [8, 8, 2000, 31, 7, 6]], dtype=np.int32)
    ref = ref_model(ids)
    d = tmp_path / "orpheus"
    d.mkdir()
    conf = dict(llama_hf_config(cfg)); conf["quantization"] = {"group_size": 64, "bits": 4}; conf["model_type"] = "llama"
    (d / "config.json").write_text(json.dumps(conf))
    keys = sorted(k for k in W if k != qname + ".weight")
    half = len(keys) // 2
    save_file_torch({k: W[k].contiguous() for k in keys[:half]}, str(d / "model-00001-of-00002.safetensors"))
    shard2 = {k: W[k].contiguous() for k in keys[half:]}
    shard2["model.layers.0.self_attn.rotary_emb.inv_freq"] = torch.ones(4)
    save_file_torch(shard2, str(d / "model-00002-of-00002.safetensors"))
    save_file({qname + ".weight": words.view(np.int32), qname + ".scales": scales, qname + ".biases": biases}, str(d / "model-quant.safetensors"))
    m = b2a.LlamaTTSModel.from_model_directory(d, max_batch=2, max_context=64)
    got = m(ids)              # stream-K partial sums land with atomics: run-to-run differences of ~1e-6 are expected
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 2e-5
    (d / "config.json").write_text("{}")
    with pytest.raises(b2a.AudioGenerationError) as e:
        b2a.LlamaTTSModel.from_model_directory(d)
    assert e.value.case == "modelNotInitialized"


def test_qwen3_talker_from_model_directory_8bit(b2a, tmp_path):
    """An 8-bit (group 64) Qwen3-TTS-style directory: "talker."-prefixed keys, one quantised Linear and the quantised codec head,
    config.json with talker_config / code_predictor_config -> the same logits and greedy frames as the model built from the
    de-quantised weights."""
    from oracle import qwen3_tts as ot
    from test_gpu_qwen3_talker import small_cfg, bf16_weights, device_model, CHAT, TTS
    cfg = small_cfg()
    W = bf16_weights(cfg, 9)
    qnames = ["model.layers.1.mlp.gate_proj", "codec_head", "code_predictor.lm_head.1"]
    quant, Wd = {}, dict(W)
    for qn in qnames:
        w32 = W[qn + ".weight"].to(torch.float32).numpy()
        words, scales, biases, q = mlx_affine_quantize(w32, 64, 8)
        quant[qn] = (words, scales, biases)
        Wd[qn + ".weight"] = torch.from_numpy((np.repeat(scales, 64, axis=1) * q + np.repeat(biases, 64, axis=1)).astype(np.float32)).to(torch.bfloat16).to(torch.float64)
    ref_model = device_model(b2a, cfg, Wd, max_batch=2, max_context=64)
    d = tmp_path / "qwen3"
    d.mkdir()
    cp = cfg.code_predictor
    conf = {"model_type": "qwen3_tts", "quantization": {"group_size": 64, "bits": 8},
            "talker_config": {"vocab_size": cfg.vocab_size, "hidden_size": cfg.hidden_size, "intermediate_size": cfg.intermediate_size,
                              "num_hidden_layers": cfg.num_hidden_layers, "num_attention_heads": cfg.num_attention_heads,
                              "num_key_value_heads": cfg.num_key_value_heads, "head_dim": cfg.head_dim, "num_code_groups": cfg.num_code_groups,
                              "text_hidden_size": cfg.text_hidden_size, "text_vocab_size": cfg.text_vocab_size,
                              "code_predictor_config": {"vocab_size": cp.vocab_size, "hidden_size": cp.hidden_size, "intermediate_size": cp.intermediate_size,
                                                        "num_hidden_layers": cp.num_hidden_layers, "num_attention_heads": cp.num_attention_heads,
                                                        "num_key_value_heads": cp.num_key_value_heads, "head_dim": cp.head_dim,
                                                        "num_code_groups": cp.num_code_groups}}}
    (d / "config.json").write_text(json.dumps(conf))
    plain = {"talker." + k: v.to(torch.bfloat16).contiguous() for k, v in W.items() if k[:-len(".weight")] not in qnames}
    plain["speaker_encoder.fc.weight"] = torch.ones(2, 2)
    save_file_torch(plain, str(d / "model.safetensors"))
    qd = {}
    for qn, (words, scales, biases) in quant.items():
        qd["talker." + qn + ".weight"], qd["talker." + qn + ".scales"], qd["talker." + qn + ".biases"] = words.view(np.int32), scales, biases
    save_file(qd, str(d / "model-quant.safetensors"))
    m = b2a.Qwen3TTSTalker.from_model_directory(d, max_batch=2, max_context=64)
    assert m.config.hidden_size == cfg.hidden_size and m.config.code_predictor.num_hidden_layers == cp.num_hidden_layers
    ri, rt, rp = ot.prepare_generation_inputs(cfg, Wd, CHAT, **TTS, language_id=2160)
    x = ri.numpy().astype(np.float32)
    (lg, hid), (rl, rh) = m(x), ref_model(x)
    assert np.linalg.norm(lg - rl) / np.linalg.norm(rl) < 2e-5 and np.linalg.norm(hid - rh) / np.linalg.norm(rh) < 2e-5
    P = b2a.Qwen3GenerateParameters(max_tokens=5, temperature=0.0, repetition_penalty=1.05, mask_eos=True)
    a, _ = m.generate_codes(x, [rt[0].numpy()], rp[0, 0].numpy(), P)
    b, _ = ref_model.generate_codes(x, [rt[0].numpy()], rp[0, 0].numpy(), P)
    assert np.array_equal(a[0], b[0])
    (d / "config.json").write_text("[]")
    with pytest.raises(b2a.AudioGenerationError):
        b2a.Qwen3TTSTalker.from_model_directory(d)
