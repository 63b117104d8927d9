"""Runs one SNAC config-2 decode (for ncu): python tools/profile_snac.py [B] [T]"""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mlx_audio_swift_b200 as m
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
codec = m.SNAC(weights=m.SNAC.random_init_weights(1234))
rng = np.random.default_rng(2)
codes = [torch.from_numpy(rng.integers(0, 4096, size=(B, T // s), dtype=np.int32)).cuda() for s in (4, 2, 1)]
wave = torch.empty((B, 1, T * 512), device="cuda")
for _ in range(2):
    codec.decode_dev(codes, wave, seed=1, stream=codec.stream)
torch.cuda.synchronize()
print("ok", float(wave.abs().mean()))
