"""GPU diagnostic (not a test): per-stage residual-stream error of the CUDA Llama step vs the oracle."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mlx_audio_swift_b200 as m  # noqa: E402
from oracle import llama as ol  # noqa: E402

cfg = ol.LlamaConfig(hidden_size=256, num_hidden_layers=2, intermediate_size=512, num_attention_heads=2,
                     num_key_value_heads=1, head_dim=128, vocab_size=2048)
W = ol.init_weights(cfg, 1234, std=0.08)
hf = dict(hidden_size=256, num_hidden_layers=2, intermediate_size=512, num_attention_heads=2, num_key_value_heads=1,
          head_dim=128, vocab_size=2048, rms_norm_eps=1e-5, rope_theta=500000.0, tie_word_embeddings=True)
tts = m.LlamaTTSModel(hf, W, max_batch=8, max_context=256)
ids = np.random.default_rng(3).integers(0, 2048, size=(2, 12)).astype(np.int32)


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


for L in (1, 2, 12):
    tts.debug_trace(True)
    lg = tts(ids[:, :L])
    tr = tts.debug_trace(True, batch=2, read=True)
    for ra in (True, False):
        t = []
        ref = ol.LlamaOracle(cfg, W, round_acts=ra).forward(torch.as_tensor(ids[:, :L], dtype=torch.long), trace=t).numpy()
        errs = [rel(tr[i], t[i].numpy()) for i in range(len(t))]
        print(f"L={L} round_acts={ra}: logits rel {rel(lg, ref):.2e} last-pos {rel(lg[:, -1], ref[:, -1]):.2e}; "
              f"trace rel errs {['%.1e' % e for e in errs]}")
