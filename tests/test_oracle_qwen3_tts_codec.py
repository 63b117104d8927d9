"""Oracle for the Qwen3-TTS speech-tokenizer decoder (row N1, oracle only): building blocks pinned against the identically
structured ``transformers`` Qwen3-Omni Code2Wav modules and torch conv primitives; streaming == full decode.  CPU only."""

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import qwen3_tts_codec as oc

hf = pytest.importorskip("transformers.models.qwen3_omni_moe.modeling_qwen3_omni_moe")
hfc = pytest.importorskip("transformers.models.qwen3_omni_moe.configuration_qwen3_omni_moe")


def _to_mlx_conv(conv: torch.nn.Conv1d, prefix):
    return {prefix + ".conv.weight": conv.weight.detach().permute(0, 2, 1).contiguous(), prefix + ".conv.bias": conv.bias.detach()}


@pytest.mark.parametrize("k,stride,dil,T", [(7, 1, 1, 19), (7, 1, 9, 40), (3, 1, 1, 5), (1, 1, 1, 7), (4, 2, 1, 11), (5, 3, 2, 23)])
def test_causal_conv_matches_hf(k, stride, dil, T):
    torch.manual_seed(k * 100 + T)
    m = hf.Qwen3OmniMoeCausalConvNet(6, 10, k, dilation=dil, stride=stride).double()
    x = torch.randn(2, 6, T, dtype=torch.float64)
    mine = oc.CausalConv(_to_mlx_conv(m.conv, "c"), "c", stride=stride, dilation=dil)(x)
    ref = m(x)
    assert mine.shape == ref.shape and (mine - ref).abs().max() < 1e-12


def test_depthwise_causal_conv_and_convnext_match_hf():
    torch.manual_seed(3)
    m = hf.Qwen3OmniMoeConvNeXtBlock(12).double()
    with torch.no_grad():
        m.gamma.copy_(torch.randn(12) * 0.5)
    W = {"b.dwconv.conv.weight": m.dwconv.conv.weight.detach().permute(0, 2, 1).contiguous(),     # [C, 1, 7] -> [C, 7, 1]
         "b.dwconv.conv.bias": m.dwconv.conv.bias.detach(), "b.norm.weight": m.norm.weight.detach(), "b.norm.bias": m.norm.bias.detach(),
         "b.pwconv1.weight": m.pwconv1.weight.detach(), "b.pwconv1.bias": m.pwconv1.bias.detach(),
         "b.pwconv2.weight": m.pwconv2.weight.detach(), "b.pwconv2.bias": m.pwconv2.bias.detach(), "b.gamma": m.gamma.detach()}
    x = torch.randn(2, 12, 17, dtype=torch.float64)
    assert (oc.ConvNeXt(W, "b")(x) - m(x)).abs().max() < 1e-12


def test_snake_beta_and_residual_unit_match_hf():
    torch.manual_seed(5)
    m = hf.Qwen3OmniMoeCode2WavDecoderResidualUnit(8, dilation=3).double()
    with torch.no_grad():
        for a in (m.act1, m.act2):
            a.alpha.copy_(torch.randn(8) * 0.4)
            a.beta.copy_(torch.randn(8) * 0.4)
    W = {"u.act1.alpha": m.act1.alpha.detach(), "u.act1.beta": m.act1.beta.detach(), "u.act2.alpha": m.act2.alpha.detach(),
         "u.act2.beta": m.act2.beta.detach(), **_to_mlx_conv(m.conv1.conv, "u.conv1"), **_to_mlx_conv(m.conv2.conv, "u.conv2")}
    x = torch.randn(2, 8, 50, dtype=torch.float64)
    assert (oc.snake_beta(x, m.act1.alpha.detach(), m.act1.beta.detach()) - m.act1(x)).abs().max() < 1e-13
    assert (oc.ResidualUnit(W, "u", 3)(x) - m(x)).abs().max() < 1e-12


@pytest.mark.parametrize("k,stride", [(2, 2), (16, 8), (6, 3), (4, 2)])
def test_transposed_conv_matches_torch_and_trims_right(k, stride):
    torch.manual_seed(k)
    w_t = torch.randn(5, 7, k, dtype=torch.float64)                 # torch layout [in, out, k]
    b = torch.randn(7, dtype=torch.float64)
    x = torch.randn(2, 5, 9, dtype=torch.float64)
    ref = F.conv_transpose1d(x, w_t, b, stride=stride)
    w_mlx = w_t.permute(1, 2, 0).contiguous()                       # sanitize: transposed(1, 2, 0)
    assert (oc.conv_transpose1d_mlx(x, w_mlx, b, stride) - ref).abs().max() < 1e-12
    up = oc.BlockUpsample({"p.conv.weight": w_mlx, "p.conv.bias": b}, "p", stride)
    y = up(x)
    assert y.shape[-1] == 9 * stride and torch.equal(y, oc.conv_transpose1d_mlx(x, w_mlx, b, stride)[:, :, : 9 * stride])


def test_transformer_layer_matches_hf_code2wav_layer():
    cfg = oc.tiny_config(num_hidden_layers=1)
    W = oc.init_weights(cfg, 11)
    hcfg = hfc.Qwen3OmniMoeCode2WavConfig(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size, num_hidden_layers=1,
                                          num_attention_heads=cfg.num_attention_heads, num_key_value_heads=cfg.num_key_value_heads,
                                          rms_norm_eps=cfg.rms_norm_eps, layer_scale_initial_scale=cfg.layer_scale_initial_scale,
                                          sliding_window=1000, attention_bias=False, rope_theta=cfg.rope_theta)
    hcfg.head_dim = cfg.head_dim
    hcfg._attn_implementation = "eager"
    layer = hf.Qwen3OmniMoeCode2WavTransformerLayer(hcfg, 0).double()
    sd = {k[len("pre_transformer.layers.0."):]: v.double() for k, v in W.items() if k.startswith("pre_transformer.layers.0.")}
    missing, unexpected = layer.load_state_dict(sd, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    T = 9
    x = torch.randn(2, T, cfg.hidden_size, dtype=torch.float64)
    cos, sin = oc.rope_cos_sin(torch.arange(T), cfg.head_dim, cfg.rope_theta)
    mask = torch.triu(torch.full((T, T), -1e9, dtype=torch.float64), 1)
    ref = layer(x, attention_mask=mask[None, None], position_embeddings=(cos[None], sin[None]))
    ref = ref[0] if isinstance(ref, tuple) else ref
    mine = oc.PreTransformer(cfg, W)._layer(0, x, cos, sin, mask, None)
    assert (mine - ref).abs().max() < 1e-6                       # HF's RMSNorm rounds through float32 internally


def test_quantizer_decode_is_sum_of_normalised_gathers_then_projection():
    cfg = oc.tiny_config()
    W = oc.init_weights(cfg, 2)
    codes = torch.from_numpy(np.random.default_rng(0).integers(0, cfg.codebook_size, (2, cfg.num_quantizers, 6)))
    codes[0, :, 0] = 0                                            # row 0 of every codebook has usage 0 -> clamp path
    q = oc.quantizer_decode(cfg, W, codes)
    assert q.shape == (2, cfg.codebook_dim, 6)
    manual = torch.zeros(2, 6, cfg.codebook_dim, dtype=torch.float64)
    for qi in range(cfg.num_quantizers):
        name, li = ("rvq_first", qi) if qi < cfg.num_semantic_quantizers else ("rvq_rest", qi - cfg.num_semantic_quantizers)
        p = f"quantizer.{name}.vq.layers.{li}.codebook"
        emb = W[p + ".embedding_sum"].double() / torch.maximum(W[p + ".cluster_usage"].double(), torch.tensor(1e-5, dtype=torch.float64))[:, None]
        manual += emb[codes[:, qi]] @ W[f"quantizer.{name}.output_proj.weight"].double()[:, 0, :].T
    assert (q.transpose(1, 2) - manual).abs().max() < 1e-10
    # only the semantic codebook given (:114-118)
    q1 = oc.quantizer_decode(cfg, W, codes[:, :1])
    assert q1.shape == q.shape and not torch.allclose(q1, q)


def test_extra_padding_whole_frames():
    # stride 1: never any right padding; stride 2, k 4: pads odd lengths up to a whole frame
    assert all(oc.extra_padding(n, 7, 1) == 0 for n in range(1, 40))
    assert [oc.extra_padding(n, 4, 2) for n in (4, 5, 6, 7)] == [0, 1, 0, 1]


def test_full_decode_shape_range_and_defaults():
    cfg = oc.tiny_config()
    W = oc.init_weights(cfg, 4)
    assert cfg.total_upsample == 4 * 3 * 2 * 2 * 2 * 2 and oc.TokenizerDecoderConfig().total_upsample == 1920
    codes = np.random.default_rng(1).integers(0, cfg.codebook_size, (2, cfg.num_quantizers, 7))
    y = oc.SpeechTokenizerDecoder(cfg, W)(codes)
    assert y.shape == (2, 1, 7 * cfg.total_upsample) and y.dtype == torch.float64
    assert y.abs().max() <= 1.0 and y.abs().max() > 1e-3
    # causality: the first t frames of audio depend only on the first t code frames
    y5 = oc.SpeechTokenizerDecoder(cfg, W)(codes[:, :, :5])
    assert (y[:, :, : 5 * cfg.total_upsample] - y5).abs().max() < 1e-9


def _zero_upsample_bias(W):
    return {k: (torch.zeros_like(v) if k.endswith("block.1.conv.bias") else v) for k, v in W.items()}


@pytest.mark.parametrize("chunks", [[11], [1] * 11, [3, 1, 5, 2], [4, 7]])
def test_streaming_step_equals_full_decode_when_upsample_bias_is_zero(chunks):
    cfg = oc.tiny_config()
    W = _zero_upsample_bias(oc.init_weights(cfg, 6))
    codes = np.random.default_rng(2).integers(0, cfg.codebook_size, (2, cfg.num_quantizers, sum(chunks)))
    d = oc.SpeechTokenizerDecoder(cfg, W)
    full = d(codes)
    assert full.abs().max() < 1.0                                 # not saturated: the comparison sees the whole signal
    d.reset_streaming_state()
    parts, s = [], 0
    for n in chunks:
        parts.append(d.streaming_step(codes[:, :, s: s + n]))
        assert parts[-1].shape[-1] == n * cfg.total_upsample
        s += n
    assert (torch.cat(parts, dim=-1) - full).abs().max() < 1e-9
    # reset really resets: a second pass reproduces the first
    d.reset_streaming_state()
    assert (d.streaming_step(codes[:, :, : chunks[0]]) - parts[0]).abs().max() < 1e-12


def test_streaming_upsample_counts_the_bias_twice_after_each_chunk_boundary():
    """DecoderBlockUpsample.step (:548-573) as written: carried tail (with bias) + new head (with bias)."""
    torch.manual_seed(0)
    r, T = 4, 9
    w = torch.randn(6, 2 * r, 5, dtype=torch.float64)
    b = torch.randn(6, dtype=torch.float64)
    up = oc.BlockUpsample({"p.conv.weight": w, "p.conv.bias": b}, "p", r)
    x = torch.randn(1, 5, T, dtype=torch.float64)
    full = up(x)
    parts = [up.step(x[:, :, 0:3]), up.step(x[:, :, 3:4]), up.step(x[:, :, 4:9])]
    got = torch.cat(parts, dim=-1)
    expect = full.clone()
    for boundary in (3, 4):                                        # r samples after each interior boundary get + bias
        expect[:, :, boundary * r: (boundary + 1) * r] += b[None, :, None]
    assert got.shape == full.shape and (got - expect).abs().max() < 1e-12
    # and the full model with real biases: a single chunk is exact, two chunks are not
    cfg = oc.tiny_config()
    W = oc.init_weights(cfg, 6)
    codes = np.random.default_rng(2).integers(0, cfg.codebook_size, (1, cfg.num_quantizers, 6))
    d = oc.SpeechTokenizerDecoder(cfg, W)
    full = d(codes)
    d.reset_streaming_state()
    assert (d.streaming_step(codes) - full).abs().max() < 1e-9
    d.reset_streaming_state()
    two = torch.cat([d.streaming_step(codes[:, :, :3]), d.streaming_step(codes[:, :, 3:])], dim=-1)
    assert (two - full)[:, :, : 3 * cfg.total_upsample].abs().max() < 1e-9 and (two - full).abs().max() > 1e-4


def test_default_geometry_decodes_12p5_hz_codes_to_24_khz():
    """The shipped geometry (Qwen3TTSConfig.swift:358-385): 16 codebooks of 2048 x 256, 1920 samples per code frame."""
    cfg = oc.TokenizerDecoderConfig()
    W = oc.init_weights(cfg, 1)
    assert W["decoder.1.block.1.conv.weight"].shape == (768, 16, 1536) and W["decoder.6.conv.weight"].shape == (1, 7, 96)
    assert W["quantizer.rvq_rest.vq.layers.14.codebook.embedding_sum"].shape == (2048, 256)
    codes = np.random.default_rng(0).integers(0, 2048, (1, 16, 3))
    d = oc.SpeechTokenizerDecoder(cfg, W)
    y = d(codes)
    assert y.shape == (1, 1, 3 * 1920) and 1e-3 < y.abs().max() <= 1.0
    d.reset_streaming_state()
    assert (d.streaming_step(codes) - y).abs().max() < 1e-9


def test_chunked_decode_and_wrapper():
    cfg = oc.tiny_config()
    W = oc.init_weights(cfg, 8)
    ac = np.random.default_rng(3).integers(1, cfg.codebook_size, (2, 13, cfg.num_quantizers))       # [B, T, n_q]
    ac[1, 9:, :] = 0                                              # padding frames in row 1 (first codebook == 0)
    d = oc.SpeechTokenizerDecoder(cfg, W)
    codes = ac.transpose(0, 2, 1)
    full = d(codes)
    assert torch.equal(d.chunked_decode(codes), full)             # one chunk (13 <= 300) == plain decode
    ch = d.chunked_decode(codes, chunk_size=5, left_context_size=2)
    assert ch.shape == full.shape
    # restated by hand: chunk [5, 10) is decoded with 2 context frames and the context audio dropped
    mid = d(codes[:, :, 3:10])[:, :, 2 * cfg.total_upsample:]
    assert torch.equal(ch[:, :, 5 * cfg.total_upsample: 10 * cfg.total_upsample], mid)
    assert torch.equal(ch[:, :, : 5 * cfg.total_upsample], d(codes[:, :, :5]))
    wav, lengths = oc.decode(cfg, W, ac)
    assert wav.shape == (2, 13 * cfg.total_upsample) and lengths.tolist() == [13 * cfg.total_upsample, 9 * cfg.total_upsample]
    Wz = _zero_upsample_bias(W)
    parts = oc.streaming_decode(cfg, Wz, ac, chunk_tokens=4)
    assert [p.shape[-1] for p in parts] == [4 * cfg.total_upsample] * 3 + [cfg.total_upsample]
    assert np.abs(np.concatenate(parts, axis=-1) - oc.SpeechTokenizerDecoder(cfg, Wz)(codes)[:, 0].numpy()).max() < 1e-9
    # decodeChunk (Qwen3TTS.swift:214-231): row 0 of the streamed audio cut to the count of non-zero first codes (whole batch)
    one = oc.decode_chunk(cfg, W, ac[:1], chunk_tokens=300)
    assert one.shape == (13 * cfg.total_upsample,) and np.abs(one - full[0, 0].numpy()).max() < 1e-9
    ac1 = ac[1:2]
    assert oc.decode_chunk(cfg, W, ac1, chunk_tokens=300).shape == (9 * cfg.total_upsample,)


def test_sanitize_from_torch_layout_round_trip():
    cfg = oc.tiny_config()
    W = oc.init_weights(cfg, 9)
    # build a PyTorch-layout checkpoint with the original key names, then sanitize it back
    ckpt = {}
    for k, v in W.items():
        key = "decoder." + k
        if ".codebook." in key:
            key = key.replace(".codebook.", "._codebook.")
        if k.startswith("upsample."):
            key = key.replace(".layers.", ".")
        is_t = (k.startswith("upsample.") and k.endswith("layers.0.conv.weight")) or (k.startswith("decoder.") and k.endswith("block.1.conv.weight"))
        if v.ndim == 3:
            v = v.permute(2, 0, 1) if is_t else v.permute(0, 2, 1)   # [out,k,in] -> torch [in,out,k] / [out,in,k]
        ckpt["speech_tokenizer." + key] = v.contiguous()
    ckpt["speech_tokenizer.decoder.quantizer.rvq_first.vq.layers.0._codebook.initialized"] = torch.ones(1)
    ckpt["speech_tokenizer.encoder.encoder.layers.0.conv.weight"] = torch.zeros(4, 1, 7)
    back = oc.strip_decoder_prefix(oc.sanitize(ckpt))
    extra = {k for k in back if k.endswith(".initialized")}
    assert set(back) - extra == set(W)
    bad = [k for k in W if back[k].shape != W[k].shape or not torch.equal(back[k], W[k])]
    # the reference's layout heuristic (checkArrayShapeQwen3) is shape-based; list what it leaves untouched at this toy geometry
    assert all(oc.check_array_shape(tuple(ckpt_v.shape)) for k in bad
               for ckpt_v in [next(v for kk, v in ckpt.items() if kk.endswith(k.replace(".layers.", ".").replace(".codebook.", "._codebook.")))]), bad


def test_check_array_shape_heuristic():
    assert oc.check_array_shape((1536, 7, 1024)) and not oc.check_array_shape((1536, 1024, 7))
    assert oc.check_array_shape((1024, 7, 1)) and not oc.check_array_shape((1024, 1, 7))
    assert oc.check_array_shape((512, 1, 256)) and not oc.check_array_shape((512, 256, 1))
    assert not oc.check_array_shape((4, 4))


def test_oracle_reproduces_committed_golden():
    """tests/golden/qwen3_codec.npz (tests/golden/make_golden.py --only qwen3_codec): the fixture the gated GPU test also compares with."""
    import importlib.util
    from conftest import GOLDEN
    spec = importlib.util.spec_from_file_location("make_golden", GOLDEN / "make_golden.py")
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    g = np.load(GOLDEN / "qwen3_codec.npz")
    cfg = mg.qwen3_codec_config()
    W = oc.init_weights(cfg, 5)
    codes = np.random.default_rng(1).integers(0, cfg.codebook_size, (2, cfg.num_quantizers, 20))
    d = oc.SpeechTokenizerDecoder(cfg, W)
    full = d(codes)[:, 0].numpy()
    assert tuple(g["shape"]) == full.shape and np.abs(full[:, :64] - g["full_first"]).max() < 1e-6 and np.abs(full[:, -64:] - g["full_last"]).max() < 1e-6
    assert np.allclose(mg.stats(full), g["full_stats"], rtol=1e-6, atol=1e-9)
    d.reset_streaming_state()
    st = np.concatenate([d.streaming_step(codes[:, :, a:b])[:, 0].numpy() for a, b in ((0, 7), (7, 8), (8, 20))], axis=-1)
    up = cfg.total_upsample
    assert np.abs(st[:, 7 * up - 32: 8 * up + 32] - g["stream_boundary"]).max() < 1e-6
    assert np.abs(st - full).max() > 1e-4                          # the chunked stream is NOT the one-shot decode (bias counted twice)
