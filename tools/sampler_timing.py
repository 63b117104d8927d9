"""Decode-loop time of one Orpheus-3B batch-8 generate with greedy vs the reference's default top-p sampling (GPU diagnostic)."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mlx_audio_swift_b200 as m  # noqa: E402
from bench import ORPHEUS, make_prompts  # noqa: E402

tts = m.LlamaTTSModel.random_init(ORPHEUS, max_batch=8, max_context=640)
ids = make_prompts(0)
for name, kw in (("greedy", dict(temperature=0.0, top_p=1.0)), ("T=0.6 top_p=0.8", dict(temperature=0.6, top_p=0.8)), ("T=0.6 top_p=1.0", dict(temperature=0.6, top_p=1.0))):
    P = m.GenerateParameters(max_tokens=512, repetition_penalty=1.3, repetition_context_size=20, mask_eos=True, wrap_codes=True, **kw)
    tts.generate_batch(ids, P, decode_audio=False)
    ts = []
    for _ in range(3):
        _, _, info = tts.generate_batch(ids, P, decode_audio=False)
        ts.append(info.generate_time / 511 * 1e3)
    print(f"{name:18s} ms/step {np.median(ts):.4f}", flush=True)
