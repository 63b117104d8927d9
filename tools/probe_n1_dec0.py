"""Probe (GPU): the implicit-conv kernel alone on the REAL operands of the Qwen3-TTS decoder's widest convolutions (default geometry,
random init): decoder.0 (k7, 1024 -> 1536, K = 7168) on the oracle's upsample output, and block 0's first dilated conv (k7, 768 -> 768,
K = 5376) on the oracle's Snake output.  The reference is the float64 convolution of the SAME float32 inputs, so the error is the
kernel's own (operand split + accumulation); next to it the float64 model of the operand split alone (tools/n1_operand_split_study.py)."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import mlx_audio_swift_b200 as b2a
from test_gpu_implicit_conv import run
from oracle import qwen3_tts_codec as oc
import n1_operand_split_study as st
import diag_n1_stages as ds


def conv_ref(w, x, T, dil):
    """w [M, taps, C], x [1, H + T, C] (H = (taps - 1) * dil history frames in front) -> [1, T, M] float64"""
    y = np.zeros((1, T, w.shape[0]))
    for j in range(w.shape[1]):
        y += x[:, j * dil: j * dil + T, :] @ w[:, j, :].T
    return y


def conv_model(w, x, T, dil, mode):
    ws = [p.numpy() for p in st.parts(torch.from_numpy(w), mode)]
    xs = [p.numpy() for p in st.parts(torch.from_numpy(x), mode)]
    return sum(conv_ref(ws[i], xs[j], T, dil) for i, j in ((0, 0), (0, 1), (1, 0)))


def report(name, w, x, T, dil):
    w = w.astype(np.float32).astype(np.float64)
    x = x.astype(np.float32).astype(np.float64)
    for taps in sorted({1, 3, w.shape[1]}):
        ww = np.ascontiguousarray(w[:, w.shape[1] - taps:, :])
        xx = np.ascontiguousarray(x[:, (w.shape[1] - taps) * dil:, :])
        ref = conv_ref(ww, xx, T, dil)
        rms = np.sqrt((ref ** 2).mean())
        line = f"{name:14s} K={taps * w.shape[2]:6d} rms(ref) {rms:8.3f}"
        for f16 in (0, 1):
            xo, _ = run(b2a, ww, xx, T, dil=dil, want_hl=False, fp16=f16)
            e = xo - ref
            line += f" | gpu {'f16 ' if f16 else 'bf16'} rms err/rms {np.sqrt((e ** 2).mean()) / rms:.2e} shrink {(e * ref).sum() / (ref ** 2).sum():+.2e}"
        for mode in ("bf16x2", "f16x2"):
            e = conv_model(ww, xx, T, dil, mode) - ref
            line += f" | model {mode} {np.sqrt((e ** 2).mean()) / rms:.2e}"
        print(line, flush=True)


def main():
    cfg = oc.TokenizerDecoderConfig()
    W = oc.init_weights(cfg, 1)
    T = 3
    codes = np.random.default_rng(0).integers(0, cfg.codebook_size, (1, cfg.num_quantizers, T))
    stg = ds.oracle_stages(cfg, W, codes)
    h2 = stg[2]                                                   # [1, 12, 1024] upsample output
    w0 = W["decoder.0.conv.weight"].double().numpy()              # [1536, 7, 1024]
    x0 = np.concatenate([np.zeros((1, 6, h2.shape[2])), h2], axis=1)
    report("decoder.0", w0, x0, h2.shape[1], 1)
    # block 0, unit 0, conv1: input = snake(act1)(stage 10), dilation 1
    h10 = torch.from_numpy(stg[10]).transpose(1, 2)
    a = oc.snake_beta(h10, oc._w(W, "decoder.1.block.2.act1.alpha"), oc._w(W, "decoder.1.block.2.act1.beta")).transpose(1, 2).numpy()
    w1 = W["decoder.1.block.2.conv1.conv.weight"].double().numpy()
    x1 = np.concatenate([np.zeros((1, 6, a.shape[2])), a], axis=1)
    report("block0.u0.c1", w1, x1, a.shape[1], 1)
    # synthetic: same shapes, N(0, 1) data
    rng = np.random.default_rng(5)
    report("normal data", rng.standard_normal(w0.shape) / np.sqrt(7168), np.concatenate([np.zeros((1, 6, 1024)), rng.standard_normal((1, 12, 1024))], axis=1), 12, 1)


if __name__ == "__main__":
    main()
