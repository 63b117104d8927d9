"""Runs the Orpheus-3B batched prefill (8 x 64-token prompts) + a few decode steps for ncu: python tools/profile_prefill.py [tokens]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mlx_audio_swift_b200 as m  # noqa: E402
from bench import ORPHEUS, make_prompts  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
tts = m.LlamaTTSModel.random_init(ORPHEUS, max_batch=8, max_context=640)
ids = make_prompts(0)
P = m.GenerateParameters(max_tokens=n, temperature=0.6, top_p=0.8, mask_eos=True, wrap_codes=True)
for _ in range(2):
    toks, _, info = tts.generate_batch(ids, P, decode_audio=False)
print("prefill_time", info.prefill_time, "generate_time", info.generate_time)
