"""Diagnostic (GPU): Qwen3-TTS speech-tokenizer decoder vs the oracle at the default geometry, per frame, bf16 vs fp16 operand pairs
(run once per mode: B2A_ST_FP16=0/1 is read when the handle is created)."""
import os
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import importlib
from oracle import qwen3_tts_codec as oc
codec = importlib.import_module("mlx_audio_swift_b200.qwen3_tts_codec")


def make(cfg, W, **kw):
    c = codec.Qwen3TTSTokenizerDecoderConfig.from_dict({k: getattr(cfg, k) for k in cfg.__dataclass_fields__})
    return codec.Qwen3TTSSpeechTokenizerDecoder(c, weights={k: v.numpy() for k, v in W.items()}, **kw)


def run(name, cfg, T, seed=1):
    W = oc.init_weights(cfg, seed)
    m = make(cfg, W)
    codes = np.random.default_rng(0).integers(0, cfg.codebook_size, (1, cfg.num_quantizers, T))
    ref = oc.SpeechTokenizerDecoder(cfg, W)(codes).numpy()[0, 0]
    y = m(codes)[0, 0]
    up = cfg.total_upsample
    e = np.abs(y - ref) / np.abs(ref).max()
    per = [float(e[f * up:(f + 1) * up].max()) for f in range(T)]
    print(f"fp16={os.environ.get('B2A_ST_FP16', '1')} {name:24s} T={T} max {e.max():.2e} per-frame {['%.1e' % p for p in per]}", flush=True)


D = oc.TokenizerDecoderConfig
for T in (1, 3, 6):
    run("default", D(), T)
run("mid", oc.mid_config(), 6)
