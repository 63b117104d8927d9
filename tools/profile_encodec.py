"""Runs one Encodec-24kHz decode (for ncu): python tools/profile_encodec.py [B] [T]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mlx_audio_swift_b200 as m  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 750
cfg = m.EncodecConfig()
codec = m.Encodec(cfg, weights=m.Encodec.random_init_weights(cfg, 1234, n_codebooks=8))
codes = torch.from_numpy(np.random.default_rng(2).integers(0, 1024, size=(1, B, 8, T), dtype=np.int32)).cuda()
wave = torch.empty((B, T * 320, 1), device="cuda")
for _ in range(2):
    codec.decode_dev(codes, wave, stream=codec.stream)
torch.cuda.synchronize()
print("ok", float(wave.abs().mean()))
