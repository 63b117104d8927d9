"""Diagnostic (GPU): Qwen3 talker forward vs oracle for growing prompt lengths (L = 1 isolates everything but q/k)."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import mlx_audio_swift_b200 as b2a
from oracle import qwen3_tts as ot
from test_gpu_qwen3_talker import small_cfg, bf16_weights, device_model, CHAT, TTS
from conftest import rel_err

cfg = small_cfg()
W = bf16_weights(cfg, 3)
m = device_model(b2a, cfg, W, max_batch=4, max_context=128)
ri, _, _ = ot.prepare_generation_inputs(cfg, W, CHAT, **TTS, language_id=2160)
for L in (1, 2, 3, 5, ri.shape[1]):
    x = ri[:, :L]
    lg, hid = m(x.numpy().astype(np.float32))
    rl, rh = ot.Talker(cfg, W)(x, None)
    print(f"L={L} logits {rel_err(lg[0], rl[0, -1].numpy()):.3e} hidden {rel_err(hid[0], rh[0, -1].numpy()):.3e}", flush=True)
# unit gains on q/k norm
W1 = dict(W)
for k in W1:
    if k.endswith("q_norm.weight") or k.endswith("k_norm.weight"):
        W1[k] = torch.ones_like(W1[k])
m1 = device_model(b2a, cfg, W1, max_batch=4, max_context=128)
lg, hid = m1(ri.numpy().astype(np.float32))
rl, rh = ot.Talker(cfg, W1)(ri, None)
print(f"unit qk gains: logits {rel_err(lg[0], rl[0, -1].numpy()):.3e} hidden {rel_err(hid[0], rh[0, -1].numpy()):.3e}")
