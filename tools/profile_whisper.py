"""Runs one Whisper-base config-3 transcription (for ncu): python tools/profile_whisper.py [B] [steps]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mlx_audio_swift_b200 as m  # noqa: E402
from tools.bench_kernels import synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cfg = dict(vocab_size=51865, num_mel_bins=80, d_model=512, encoder_layers=6, encoder_attention_heads=8, encoder_ffn_dim=2048,
           max_source_positions=1500, decoder_layers=6, decoder_attention_heads=8, decoder_ffn_dim=2048, max_target_positions=448)
wm = m.WhisperModel.random_init(cfg, max_batch=B)
x = torch.from_numpy(np.stack([synth(480000, i) for i in range(B)])).cuda()
P = m.STTGenerateParameters(max_tokens=steps, mask_eot=True)
toks = np.zeros((B, steps), dtype=np.int32)
nt = np.zeros(B, dtype=np.int32)
for _ in range(2):
    o = wm.generate_dev(x, P, toks, nt)
torch.cuda.synchronize()
print("ok", o.encode_time, o.decode_time)
