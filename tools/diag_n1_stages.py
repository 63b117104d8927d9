"""Diagnostic (GPU): where does the Qwen3-TTS speech-tokenizer decoder leave the float64 oracle?  Compares the fp32 activation tensor
after every stage (b2a_speech_tokenizer_debug_stage) with the oracle's, relative to the oracle tensor's max.

    B2A_ST_FP16=0|1 python tools/diag_n1_stages.py [frames] [default|mid]
"""
import ctypes as C
import importlib
import os
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import qwen3_tts_codec as oc
codec = importlib.import_module("mlx_audio_swift_b200.qwen3_tts_codec")
_ffi = importlib.import_module("mlx_audio_swift_b200._ffi")


def oracle_stages(cfg, W, codes):
    """{stage id: [B, T, C] float64} following SpeechTokenizerDecoder.__call__"""
    d = oc.SpeechTokenizerDecoder(cfg, W)
    out = {}
    h = oc.quantizer_decode(cfg, W, torch.as_tensor(codes))
    h = d.pre_conv(h)
    pt = d.pre_transformer
    c = cfg
    x = pt._lin(pt.p + ".input_proj", h.transpose(1, 2))
    T = x.shape[1]
    cos, sin = oc.rope_cos_sin(torch.arange(0, T), c.head_dim, c.rope_theta)
    mask = None
    if T > 1:
        rows = torch.arange(0, T)[:, None]
        mask = torch.where(torch.arange(T)[None, :] > rows, torch.tensor(-1e9, dtype=oc.DT), torch.tensor(0.0, dtype=oc.DT))
    for i in range(c.num_hidden_layers):
        x = pt._layer(i, x, cos, sin, mask, None)
    out[0] = x.clone()
    h = pt._lin(pt.p + ".output_proj", oc.rms_norm(x, oc._w(W, pt.p + ".norm.weight"), c.rms_norm_eps)).transpose(1, 2)
    for i, u in enumerate(d.upsample):
        h = u(h)
        out[1 + i] = h.transpose(1, 2).clone()
    h = d.init_conv(h)
    for b, blk in enumerate(d.blocks):
        h = blk.up(oc.snake_beta(h, *blk.snake))
        out[10 + 4 * b] = h.transpose(1, 2).clone()
        for j, u in enumerate(blk.units):
            h = u(h)
            out[11 + 4 * b + j] = h.transpose(1, 2).clone()
    return {k: v.numpy() for k, v in out.items()}


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    geom = sys.argv[2] if len(sys.argv) > 2 else "default"
    cfg = oc.TokenizerDecoderConfig() if geom == "default" else oc.mid_config()
    W = oc.init_weights(cfg, 1)
    codes = np.random.default_rng(0).integers(0, cfg.codebook_size, (1, cfg.num_quantizers, T))
    ref = oracle_stages(cfg, W, codes)
    c = codec.Qwen3TTSTokenizerDecoderConfig.from_dict({k: getattr(cfg, k) for k in cfg.__dataclass_fields__})
    m = codec.Qwen3TTSSpeechTokenizerDecoder(c, weights={k: v.numpy() for k, v in W.items()})
    lib = _ffi.lib()
    names = {0: "transformer"}
    for k in sorted(ref):
        if k == 0:
            continue
        names[k] = f"upsample {k - 1}" if k < 10 else (f"block {(k - 10) // 4} convT" if (k - 10) % 4 == 0 else f"block {(k - 10) // 4} unit {(k - 10) % 4 - 1}")
    for k in sorted(ref):
        _ffi.check(lib.b2a_speech_tokenizer_debug_stage(m._h, k, None, 0, None))
        m(codes)
        buf = np.empty(ref[k].size, np.float32)
        n = C.c_int64(0)
        _ffi.check(lib.b2a_speech_tokenizer_debug_stage(m._h, -1, _ffi.ptr(buf), buf.size, C.byref(n)))
        assert n.value == ref[k].size, (k, n.value, ref[k].shape)
        y = buf.reshape(ref[k].shape)
        e = np.abs(y - ref[k])
        scale = np.abs(ref[k]).max()
        rms = np.sqrt((ref[k] ** 2).mean())
        frames = ref[k].shape[1] // T
        per = ["%.1e" % (e[:, f * frames:(f + 1) * frames].max() / scale) for f in range(T)]
        print(f"fp16={os.environ.get('B2A_ST_FP16', '1')} {geom} T={T} stage {k:2d} {names[k]:16s} shape {tuple(ref[k].shape)} max|ref| {scale:9.3e} rms {rms:9.3e} "
              f"max err/max {e.max() / scale:.2e}  rms err/rms {np.sqrt((e ** 2).mean()) / rms:.2e}  per-frame {per}", flush=True)


if __name__ == "__main__":
    main()
