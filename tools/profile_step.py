"""Runs a few captured Orpheus-3B decode steps (for ncu): python tools/profile_step.py [ctx] [iters]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mlx_audio_swift_b200 as m  # noqa: E402
from bench import ORPHEUS  # noqa: E402

ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 320
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
tts = m.LlamaTTSModel.random_init(ORPHEUS, max_batch=8, max_context=640)
print("ms/step", tts.time_steps(8, ctx, iters))
