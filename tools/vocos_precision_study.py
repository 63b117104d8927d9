"""CPU study: operand scheme the Vocos decode's GEMMs need (embed conv, pwconv1/2, head projection, windowed inverse DFT as a matrix).
The float64 oracle pipeline (oracle/vocos.py) is replayed with every GEMM's operands rounded as a tensor-core kernel would.

    python tools/vocos_precision_study.py
Results: profiles/r01_snac_precision_study.md (last paragraph)."""
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import vocos as ov  # noqa: E402

T64 = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float64)


def rnd(x, kind):
    return x.to(torch.float32).to(torch.bfloat16 if kind == "bf16" else torch.float16).to(torch.float64)


def mm(x, W, mode):
    """x [.., K] @ W[M, K]^T with the operands of `mode`."""
    if mode == "exact":
        return x @ W.T
    kind, n = mode.split(" x")
    xh, wh = rnd(x, kind), rnd(W, kind)
    y = xh @ wh.T
    if n in ("2", "3"):
        y = y + rnd(x - xh, kind) @ wh.T
    if n == "3":
        y = y + xh @ rnd(W - wh, kind).T
    return y.to(torch.float32).to(torch.float64)


def decode(cfg, w, feats, mode):
    d, k = cfg.dim, cfg.input_kernel_size
    x = T64(feats)
    B, L, Cin = x.shape
    xp = F.pad(x, (0, 0, k // 2, k // 2))
    cols = torch.cat([xp[:, i:i + L] for i in range(k)], dim=-1)                     # im2col [B, L, k * Cin]
    h = mm(cols, T64(w["backbone.embed.weight"]).reshape(d, k * Cin), mode) + T64(w["backbone.embed.bias"])
    h = F.layer_norm(h, (d,), T64(w["backbone.norm.weight"]), T64(w["backbone.norm.bias"]), 1e-6)
    for l in range(cfg.num_layers):
        p = f"backbone.convnext.{l}."
        y = F.conv1d(h.transpose(1, 2), T64(w[p + "dwconv.weight"]).permute(0, 2, 1), T64(w[p + "dwconv.bias"]), padding=cfg.dw_kernel_size // 2, groups=d).transpose(1, 2)
        y = F.layer_norm(y, (d,), T64(w[p + "norm.weight"]), T64(w[p + "norm.bias"]), 1e-6)
        y = F.gelu(mm(y, T64(w[p + "pwconv1.weight"]), mode) + T64(w[p + "pwconv1.bias"]))
        y = mm(y, T64(w[p + "pwconv2.weight"]), mode) + T64(w[p + "pwconv2.bias"])
        h = h + T64(w[p + "gamma"]) * y
    h = F.layer_norm(h, (d,), T64(w["backbone.final_layer_norm.weight"]), T64(w["backbone.final_layer_norm.bias"]), 1e-6)
    hh = mm(h, T64(w["head.out.weight"]), mode) + T64(w["head.out.bias"])
    N, half = cfg.n_fft, cfg.n_fft // 2 + 1
    mag = torch.clamp(torch.exp(hh[..., :half]), max=1e2)
    spec = torch.cat([mag * torch.cos(hh[..., half:]), mag * torch.sin(hh[..., half:])], dim=-1)       # [B, L, 2 * half]
    win = ov.hann_symmetric(N)
    j, kq = torch.arange(N, dtype=torch.float64)[:, None], torch.arange(half, dtype=torch.float64)[None, :]
    ck = torch.full((half,), 2.0, dtype=torch.float64); ck[0] = 1.0; ck[-1] = 1.0
    A = torch.cat([win[:, None] * ck * torch.cos(2 * np.pi * j * kq / N) / N, -win[:, None] * ck * torch.sin(2 * np.pi * j * kq / N) / N], dim=1)
    A[:, half] = 0.0; A[:, 2 * half - 1] = 0.0                                         # the imaginary parts of DC / Nyquist do not enter irfft
    frames = mm(spec, A, mode)                                                         # windowed inverse real DFT as a matrix
    out_len = (L - 1) * cfg.hop_length + N
    audio, wsum = torch.zeros(B, out_len, dtype=torch.float64), torch.zeros(out_len, dtype=torch.float64)
    for i in range(L):
        audio[:, i * cfg.hop_length: i * cfg.hop_length + N] += frames[:, i]
        wsum[i * cfg.hop_length: i * cfg.hop_length + N] += win
    audio = audio / wsum.clamp(min=1e-300)
    return audio[:, N // 2: out_len - N // 2].numpy()


def main():
    cfg = ov.VocosConfig(num_layers=4)
    W = ov.init_weights(cfg, 7)
    f = np.random.default_rng(1).standard_normal((1, 37, cfg.input_channels)).astype(np.float32)
    ref = ov.decode(cfg, W, f)
    assert np.abs(decode(cfg, W, f, "exact") - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())     # the replay IS the oracle
    for mode in ("bf16 x3", "bf16 x1", "fp16 x2", "fp16 x1"):
        y = decode(cfg, W, f, mode)
        print(f"{mode:8s} max err / peak {np.abs(y - ref).max() / np.abs(ref).max():.2e}   rel L2 {np.linalg.norm(y - ref) / np.linalg.norm(ref):.2e}")


if __name__ == "__main__":
    main()
