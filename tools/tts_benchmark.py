"""The `--benchmark` report of the reference CLI (Sources/Tools/mlx-audio-swift-tts/App.swift:128-212) for the B200 path:
Audio duration / TTFB / RTFx / Tokens/s of one Orpheus-3B generate call (random-init weights, synthetic prompt ids -- there is no
tokenizer or checkpoint here).  The reference's Orpheus emits its audio once, at the end (LlamaTTS.swift:901-904), so its TTFB is the
whole generation; with --stream (row N2, b2a_tts_generate_stream) audio chunks are decoded by SNAC while tokens are still being
generated and TTFB is the latency of the first .audio event, as the CLI measures it (App.swift:155-170, streamingInterval 0.32 s).

    python tools/tts_benchmark.py [--batch 1] [--prompt 64] [--max-tokens 512] [--model-dir DIR] [--stream] [--interval 0.32]"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mlx_audio_swift_b200 as m  # noqa: E402
from bench import ORPHEUS, make_prompts  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--prompt", type=int, default=64)
ap.add_argument("--max-tokens", type=int, default=512)
ap.add_argument("--model-dir", default=None, help="checkpoint directory (config.json + *.safetensors); default: random init")
ap.add_argument("--stream", action="store_true", help="chunked audio emission during generation (TTFB = first audio chunk)")
ap.add_argument("--interval", type=float, default=0.32, help="streaming interval in seconds (App.swift:137)")
a = ap.parse_args()
codec = m.SNAC(weights=m.SNAC.random_init_weights(1234))
if a.model_dir:
    tts = m.LlamaTTSModel.from_model_directory(a.model_dir, snac=codec, max_batch=a.batch, max_context=a.prompt + a.max_tokens + 16)
else:
    tts = m.LlamaTTSModel.random_init(ORPHEUS, snac=codec, max_batch=a.batch, max_context=a.prompt + a.max_tokens + 16)
ids = make_prompts(0)[:a.batch, :a.prompt]
P = m.GenerateParameters(max_tokens=a.max_tokens, temperature=0.6, top_p=0.8, repetition_penalty=1.3, repetition_context_size=20,
                         mask_eos=True, wrap_codes=True)
tts.generate_batch(ids, P)                       # warm-up (graph capture, allocations)
first_token, first_audio = [], []
if a.stream:
    fpc = max(1, int(round(a.interval * 24000.0 / 2048.0)))
    tts.generate_audio_chunks(ids, P, frames_per_chunk=fpc)          # warm-up of the chunked codec shapes
    started = time.perf_counter()
    toks, chunks, info = tts.generate_audio_chunks(ids, P, frames_per_chunk=fpc,
                                                    on_audio=lambda b, x, fin: first_audio.append(time.perf_counter()) if not first_audio else None,
                                                    on_token=lambda b, step, tok: first_token.append(time.perf_counter()) if not first_token else None)
    elapsed = time.perf_counter() - started
    waves = [np.concatenate(c) if c else None for c in chunks]
else:
    started = time.perf_counter()
    toks, waves, info = tts.generate_batch(ids, P, on_token=lambda b, step, tok: first_token.append(time.perf_counter()) if not first_token else None)
    elapsed = time.perf_counter() - started
audio = sum(len(w) for w in waves if w is not None) / 24000.0
print(f"Finished generation in {elapsed:0.2f}s")
print("Benchmark:")
print(f"  Audio duration: {audio:.2f}s")
ttfb = (first_audio[0] - started) if first_audio else elapsed
print(f"  TTFB: {ttfb:.3f}s (first token after {first_token[0] - started:.3f}s)" if first_token else "  TTFB: n/a")
print(f"  RTFx: {audio / elapsed:.3f}" if audio > 0 else "  RTFx: n/a")
print(f"  Tokens/s: {info.tokens_per_second:.2f}")
