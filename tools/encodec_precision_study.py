"""CPU study: operand-splitting scheme the Encodec decoder's convolutions would need on tensor cores (they are fp32 SIMT today).

The float64 oracle decode (oracle/encodec.py, 24 kHz geometry, random-init weights) with every convolution / transposed convolution's
operands rounded as a tensor-core kernel would (bf16 or fp16, optionally hi + lo pairs), fp32 result; the LSTM stays exact.

    python tools/encodec_precision_study.py [frames]

Results: profiles/r01_encodec_precision_study.md."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import encodec as oe  # noqa: E402


def rnd(x, kind):
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    return t.to(torch.bfloat16 if kind == "bf16" else torch.float16).to(torch.float32).numpy().astype(np.float64)


def pair(x, kind):
    hi = rnd(x, kind)
    return hi, rnd(x - hi, kind)


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    cfg = oe.EncodecConfig()
    W = oe.init_weights(cfg, 7, n_codebooks=8)
    codes = np.random.default_rng(1).integers(0, 1024, size=(1, 1, 8, T))
    ref = oe.decode(cfg, W, codes)
    c0, t0 = oe.conv1d, oe.conv_transpose1d
    zero = np.zeros(1)
    for mode in ("bf16 x3", "bf16 x1", "fp16 x2", "fp16 x1"):
        kind, n = mode.split(" x")

        def products(f, x, w):
            wh, wl = pair(w, kind)
            xh, xl = pair(x, kind)
            y = f(xh, wh)
            if n in ("2", "3"):
                y = y + f(xl, wh)
            if n == "3":
                y = y + f(xh, wl)
            return y.astype(np.float32).astype(np.float64)

        def conv1d(cfg_, x, w, b, stride=1, dilation=1):
            return products(lambda xx, ww: c0(cfg_, xx, ww, np.zeros_like(b), stride, dilation), x, w) + b.astype(np.float64)

        def conv_t(cfg_, x, w, b, stride):
            return products(lambda xx, ww: t0(cfg_, xx, ww, np.zeros_like(b), stride), x, w) + b.astype(np.float64)

        oe.conv1d, oe.conv_transpose1d = conv1d, conv_t
        y = oe.decode(cfg, W, codes)
        oe.conv1d, oe.conv_transpose1d = c0, t0
        print(f"{mode:8s} max err / peak {np.abs(y - ref).max() / np.abs(ref).max():.2e}   rel L2 {np.linalg.norm(y - ref) / np.linalg.norm(ref):.2e}   {ref.shape}")


if __name__ == "__main__":
    main()
