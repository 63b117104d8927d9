"""Diagnostic (GPU): where does the Qwen3-TTS speech-tokenizer decoder leave the oracle at the default geometry?"""
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
    peak = np.abs(ref).max()
    e = np.abs(y - ref) / peak
    per = [float(e[f * up:(f + 1) * up].max()) for f in range(T)]
    first = int(np.argmax(e > 2e-4)) if (e > 2e-4).any() else -1
    print(f"{name:48s} T={T} max {e.max():.2e} argmax {int(e.argmax())} first>2e-4 at {first} per-frame {['%.1e' % p for p in per]}", flush=True)


D = oc.TokenizerDecoderConfig
for T in (1, 2, 3, 5):
    run("default", D(), T)
run("default 1 layer", D(num_hidden_layers=1), 3)
run("default rates [8,5]", D(upsample_rates=[8, 5]), 3)
run("default rates [4,3,2]", D(upsample_rates=[4, 3, 2]), 3)
run("default dim 768", D(decoder_dim=768), 3)
run("default heads 4x32 kv2", D(num_attention_heads=4, num_key_value_heads=2, head_dim=32), 3)
run("default hidden 64 inter 128", D(hidden_size=64, intermediate_size=128), 3)
run("default nq 4", D(num_quantizers=4), 3)
run("default codebook 64 dim 128 latent 128", D(codebook_size=64, codebook_dim=128, latent_dim=128), 3)
run("default layer_scale 0.3", D(layer_scale_initial_scale=0.3), 3)
run("mid", oc.mid_config(), 3)
run("mid rates [8,5,4,3]", oc.mid_config(upsample_rates=[8, 5, 4, 3]), 3)
run("mid dim 1536", oc.mid_config(decoder_dim=1536), 3)
