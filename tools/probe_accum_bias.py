"""Probe (GPU): is the tcgen05 fp32 accumulation biased?  Positive bf16-exact operands (lo halves are zero, every product is exact in
fp32), so any error against the float64 sum is the accumulator's rounding.  Prints mean signed relative error vs K."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import mlx_audio_swift_b200 as b2a
from test_gpu_implicit_conv import run

for signed in (False, True):
    for cin, taps in ((64, 1), (64, 7), (384, 7), (1536, 7)):
        rng = np.random.default_rng(cin + taps)
        w = rng.random((128, taps, cin)) + 0.5
        x = rng.random((1, taps - 1 + 64, cin)) + 0.5
        if signed:
            w *= rng.choice([-1.0, 1.0], size=w.shape); x *= rng.choice([-1.0, 1.0], size=x.shape)
        w = torch.from_numpy(w).to(torch.bfloat16).double().numpy()
        x = torch.from_numpy(x).to(torch.bfloat16).double().numpy()
        ref = np.zeros((1, 64, 128))
        for j in range(taps):
            ref += x[:, j:j + 64, :] @ w[:, j, :].T
        for f16 in (0, 1):
            xo, _ = run(b2a, w, x, 64, want_hl=False, fp16=f16)
            scale = np.abs(ref) if not signed else np.sqrt((x[:, :64] ** 2).sum(-1, keepdims=True) * (w ** 2).sum((1, 2))[None, None, :] / cin)
            rel = (xo - ref) / scale
            print(f"signed={signed} K={cin * taps:6d} f16={f16} mean signed rel err {rel.mean():+.3e}  rms {np.sqrt((rel ** 2).mean()):.3e}  max {np.abs(rel).max():.3e}", flush=True)
