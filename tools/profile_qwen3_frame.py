"""Runs a few Qwen3-TTS frames (talker step + 15 code-predictor passes, one CUDA graph per frame) for ncu:
    python tools/profile_qwen3_frame.py [rows] [frames]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mlx_audio_swift_b200 as m  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 6
cfg = m.Qwen3TalkerConfig()
talker = m.Qwen3TTSTalker.random_init(cfg, max_batch=rows, max_context=frames + 32, std=0.02, seed=77)
rng = np.random.default_rng(9)
H = cfg.hidden_size
embeds = (0.05 * rng.standard_normal((rows, 10, H))).astype(np.float32)
trailing = (0.05 * rng.standard_normal((rows, 4, H))).astype(np.float32)
pad = (0.05 * rng.standard_normal(H)).astype(np.float32)
P = m.Qwen3GenerateParameters(max_tokens=frames, temperature=0.9, top_k=50, top_p=1.0, repetition_penalty=1.05, seed=1, mask_eos=True)
codes, info = talker.generate_codes(embeds, list(trailing), pad, P)
print("frames", [len(c) for c in codes], "generate_time", info.generate_time)
