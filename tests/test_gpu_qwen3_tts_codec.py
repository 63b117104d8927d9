"""CUDA Qwen3-TTS speech-tokenizer decoder (through the C ABI) vs the oracle (row N1).

Part of the default ``-m gpu`` run since round 2.  Every structural feature of the shipped decoder is held to 1e-3 at the mid
geometry and at the shipped (default) one, the latter with both operand formats (test_default_geometry_and_errors)."""
import os

import numpy as np
import pytest

from conftest import max_rel_to_peak, rel_err
from oracle import qwen3_tts_codec as oc

pytestmark = pytest.mark.gpu
TOL = 1e-3


mid_config = oc.mid_config


def make(b2a_codec, cfg, W, **kw):
    c = b2a_codec.Qwen3TTSTokenizerDecoderConfig.from_dict({k: getattr(cfg, k) for k in cfg.__dataclass_fields__})
    return b2a_codec.Qwen3TTSSpeechTokenizerDecoder(c, weights={k: v.numpy() for k, v in W.items()}, **kw)


@pytest.fixture(scope="module")
def codec():
    import importlib
    return importlib.import_module("mlx_audio_swift_b200.qwen3_tts_codec")


def test_one_shot_decode_vs_oracle(codec):
    cfg = mid_config()
    W = oc.init_weights(cfg, 5)
    m = make(codec, cfg, W, max_batch=2)
    codes = np.random.default_rng(1).integers(0, cfg.codebook_size, (2, cfg.num_quantizers, 70))      # crosses a 64-frame tile
    ref = oc.SpeechTokenizerDecoder(cfg, W)(codes).numpy()
    y = m(codes)
    assert y.shape == ref.shape == (2, 1, 70 * cfg.total_upsample)
    assert max_rel_to_peak(y, ref) < TOL and rel_err(y, ref) < TOL, (max_rel_to_peak(y, ref), rel_err(y, ref))
    assert np.abs(m(codes[1:2]) - y[1:2]).max() < 1e-6                                             # batched == serial
    y1 = m(codes[:, :1])                                                                            # only the semantic codebook
    assert max_rel_to_peak(y1, oc.SpeechTokenizerDecoder(cfg, W)(codes[:, :1]).numpy()) < TOL


@pytest.mark.parametrize("chunks", [[1] * 5, [3, 1, 7, 2], [40, 30]])
def test_streaming_step_vs_oracle_including_the_bias_quirk(codec, chunks):
    cfg = mid_config()
    W = oc.init_weights(cfg, 6)
    m = make(codec, cfg, W, max_batch=2)
    d = oc.SpeechTokenizerDecoder(cfg, W)
    codes = np.random.default_rng(2).integers(0, cfg.codebook_size, (2, cfg.num_quantizers, sum(chunks)))
    m.reset_streaming_state(); d.reset_streaming_state()
    s = 0
    for n in chunks:
        y, ref = m.streaming_step(codes[:, :, s: s + n]), d.streaming_step(codes[:, :, s: s + n]).numpy()
        assert y.shape == ref.shape and max_rel_to_peak(y, ref) < TOL, (s, max_rel_to_peak(y, ref))
        s += n


def test_chunked_and_streaming_decode_wrappers(codec):
    cfg = mid_config()
    W = oc.init_weights(cfg, 8)
    ac = np.random.default_rng(3).integers(1, cfg.codebook_size, (2, 13, cfg.num_quantizers))
    ac[1, 9:, :] = 0
    tok = codec.Qwen3TTSSpeechTokenizer(codec.Qwen3TTSTokenizerDecoderConfig.from_dict({k: getattr(cfg, k) for k in cfg.__dataclass_fields__}),
                                        weights={k: v.numpy() for k, v in W.items()}, decode_upsample_rate=cfg.total_upsample, max_batch=2)
    d = oc.SpeechTokenizerDecoder(cfg, W)
    codes = ac.transpose(0, 2, 1)
    ch = tok.decoder.chunked_decode(codes, chunk_size=5, left_context_size=2)
    assert max_rel_to_peak(ch, d.chunked_decode(codes, 5, 2).numpy()) < TOL
    wav, lengths = tok.decode(ac)
    ref_wav, ref_len = oc.decode(cfg, W, ac)
    assert lengths.tolist() == ref_len.tolist() and max_rel_to_peak(wav, ref_wav) < TOL
    parts = tok.streaming_decode(ac, chunk_tokens=4)
    refs = oc.streaming_decode(cfg, W, ac, chunk_tokens=4)
    assert [p.shape for p in parts] == [r.shape for r in refs]
    assert max(max_rel_to_peak(p, r) for p, r in zip(parts, refs)) < TOL
    assert max_rel_to_peak(tok.decode_chunk(ac[:1], 300), oc.decode_chunk(cfg, W, ac[:1], 300)) < TOL


def test_default_geometry_and_errors(b2a, codec):
    cfg = oc.TokenizerDecoderConfig()
    W = oc.init_weights(cfg, 1)
    m = make(codec, cfg, W)
    assert m.total_upsample == 1920
    codes = np.random.default_rng(0).integers(0, 2048, (1, 16, 6))
    ref = oc.SpeechTokenizerDecoder(cfg, W)(codes).numpy()
    # Round 2 sat at 2.4e-3 here: tcgen05's fp32 accumulation truncates, a multiplicative bias of -1.8e-9 * K per convolution (-1.2e-5 at
    # K = 7168, tools/probe_n1_dec0.py) that this stack amplifies ~70x.  With the contraction accumulated in segments of 256 and the
    # segments added in registers (implicit_conv.cuh, Args::seg_kb) the measured error at 6 frames is 2.7e-4 of the peak with fp16 operand
    # pairs (the default) and 7.0e-4 with bf16 pairs (profiles/r02_n1_decoder_numerics.md).
    assert max_rel_to_peak(m(codes), ref) < TOL
    os.environ["B2A_ST_FP16"] = "0"                              # read when a handle is created
    try:
        assert max_rel_to_peak(make(codec, cfg, W)(codes), ref) < TOL
    finally:
        del os.environ["B2A_ST_FP16"]
    with pytest.raises(b2a.AudioGenerationError) as e:
        m(np.zeros((2, 16, 3), np.int32))                         # batch > max_batch
    assert e.value.case == "invalidInput"
    W2 = dict(W); W2.pop("pre_conv.conv.bias")
    with pytest.raises(b2a.AudioGenerationError) as e:
        make(codec, cfg, W2)
    assert e.value.case == "modelNotInitialized"


def test_from_model_directory(codec, tmp_path):
    """loadSpeechTokenizer: a PyTorch-layout safetensors checkpoint + config.json through the library's own sanitize."""
    import json
    from safetensors.torch import save_file
    from test_qwen3_tts_codec_host import torch_layout_checkpoint
    # every k = 1 conv needs > 64 input channels here: the reference's layout heuristic (checkArrayShapeQwen3) reads a PyTorch
    # [out, <= 64, 1] weight as "already MLX" and would leave it untransposed (true of the reference itself, not only of this port)
    cfg = mid_config(codebook_dim=144, decoder_dim=512, upsample_rates=[4, 3])     # 512 -> 256 -> 128 output channels (the output conv holds <= 128)
    W = oc.init_weights(cfg, 12)
    d = tmp_path / "speech_tokenizer"
    d.mkdir()
    save_file({k: v.contiguous() for k, v in torch_layout_checkpoint(W).items()}, str(d / "model.safetensors"))
    keys = ("latent_dim", "codebook_dim", "codebook_size", "decoder_dim", "hidden_size", "intermediate_size", "head_dim", "num_attention_heads",
            "num_key_value_heads", "num_hidden_layers", "num_quantizers", "num_semantic_quantizers", "upsample_rates", "upsampling_ratios")
    (d / "config.json").write_text(json.dumps({"decode_upsample_rate": cfg.total_upsample, "decoder_config": {k: getattr(cfg, k) for k in keys}}))
    m = codec.Qwen3TTSSpeechTokenizerDecoder.from_model_directory(d)
    assert m.total_upsample == cfg.total_upsample == m.decode_upsample_rate
    codes = np.random.default_rng(5).integers(0, cfg.codebook_size, (1, cfg.num_quantizers, 9))
    Ws = oc.strip_decoder_prefix(oc.sanitize(torch_layout_checkpoint(W)))
    assert all(tuple(Ws[k].shape) == tuple(W[k].shape) for k in W)
    assert max_rel_to_peak(m(codes), oc.SpeechTokenizerDecoder(cfg, W)(codes).numpy()) < TOL


def test_decode_vs_committed_golden(codec):
    """tests/golden/qwen3_codec.npz: first / last 64 samples + stats of the one-shot decode, and the samples around two chunk
    boundaries of a streamed decode (where the reference counts the transposed-conv bias twice)."""
    import importlib.util
    from conftest import GOLDEN
    spec = importlib.util.spec_from_file_location("make_golden", GOLDEN / "make_golden.py")
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    g = np.load(GOLDEN / "qwen3_codec.npz")
    cfg = mg.qwen3_codec_config()
    W = oc.init_weights(cfg, 5)
    m = make(codec, cfg, W, max_batch=2)
    codes = np.random.default_rng(1).integers(0, cfg.codebook_size, (2, cfg.num_quantizers, 20))
    peak = max(abs(g["full_stats"][2]), abs(g["full_stats"][3]))
    y = m(codes)[:, 0]
    assert y.shape == tuple(g["shape"]) and np.abs(y[:, :64] - g["full_first"]).max() < TOL * peak and np.abs(y[:, -64:] - g["full_last"]).max() < TOL * peak
    assert np.abs(mg.stats(y) - g["full_stats"]).max() < TOL * peak
    m.reset_streaming_state()
    st = np.concatenate([m.streaming_step(codes[:, :, a:b])[:, 0] for a, b in ((0, 7), (7, 8), (8, 20))], axis=-1)
    up = cfg.total_upsample
    assert np.abs(st[:, 7 * up - 32: 8 * up + 32] - g["stream_boundary"]).max() < TOL * peak
