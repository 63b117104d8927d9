"""CPU study: what operand precision does a tensor-core Whisper ENCODER attention need to keep the encoder output inside 1e-3?

The fp32 oracle (oracle/whisper.py) is re-run with `softmax(q k^T / sqrt(d)) v` of the ENCODER emulated as a tensor-core kernel would
compute it: q, k, v (and the probabilities p before the second product) rounded to bf16 or fp16, optionally q / k / p as hi + lo pairs.

    python tools/whisper_attention_precision_study.py [tiny|base]

Prints the relative L2 error of the encoder output against the exact run, and whether the 12 greedy tokens of the golden clip change.
Results: profiles/r01_whisper_attention_precision_study.md."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import dsp, whisper as ow  # noqa: E402


def rnd(x, kind):
    return x.to(torch.bfloat16 if kind == "bf16" else torch.float16).to(torch.float32)


def make_sdpa(mode):
    kind, split = mode.split("-") if "-" in mode else (mode, "")

    def pair(x):
        hi = rnd(x, kind)
        return hi, rnd(x - hi, kind)

    def sdpa(self, q, k, v, nh, mask=None):
        hd = q.shape[-1]
        if mask is not None or mode == "exact":                      # decoder (causal) attention stays exact: the study is about the encoder
            s = (q @ k.transpose(-1, -2)) * hd ** -0.5
            if mask is not None:
                s = s + mask
            o = torch.softmax(s, dim=-1) @ v
        else:
            if split == "hilo":                                      # q, k as hi/lo (3 products), p as hi/lo x v hi (2 products)
                qh, ql = pair(q); kh, kl = pair(k)
                s = (qh @ kh.transpose(-1, -2) + qh @ kl.transpose(-1, -2) + ql @ kh.transpose(-1, -2)) * hd ** -0.5
                p = torch.softmax(s, dim=-1)
                ph, pl = pair(p); vh, vl = pair(v)
                o = ph @ vh + pl @ vh + ph @ vl
            else:
                s = (rnd(q, kind) @ rnd(k, kind).transpose(-1, -2)) * hd ** -0.5
                p = torch.softmax(s, dim=-1)
                o = rnd(p, kind) @ rnd(v, kind)
        B, _, T, _ = o.shape
        return o.transpose(1, 2).reshape(B, T, nh * hd)
    return sdpa


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    cfg = ow.WhisperConfig.tiny_test() if which == "tiny" else ow.WhisperConfig(d_model=512, encoder_layers=6, encoder_attention_heads=8, encoder_ffn_dim=2048,
                                                                                  decoder_layers=6, decoder_attention_heads=8, decoder_ffn_dim=2048)
    W = ow.init_weights(cfg, 1234)
    x = dsp.synth_audio(64000, 3)
    feats = torch.from_numpy(dsp.whisper_encoder_features(x, cfg.num_mel_bins)).float()
    orig = ow.WhisperOracle._sdpa
    ref_enc = ow.WhisperOracle(cfg, W).encode(feats)
    ref_tok = ow.transcribe_tokens(ow.WhisperOracle(cfg, W), x, ow.build_prompt_tokens(), max_tokens=12, mask_eot=True)
    for mode in ("exact", "bf16", "fp16", "bf16-hilo", "fp16-hilo"):
        ow.WhisperOracle._sdpa = make_sdpa(mode)
        enc = ow.WhisperOracle(cfg, W).encode(feats)
        tok = ow.transcribe_tokens(ow.WhisperOracle(cfg, W), x, ow.build_prompt_tokens(), max_tokens=12, mask_eot=True)
        err = float(torch.linalg.norm(enc - ref_enc) / torch.linalg.norm(ref_enc))
        print(f"{which:5s} {mode:10s} encoder rel L2 {err:.2e}   max/peak {float((enc - ref_enc).abs().max() / ref_enc.abs().max()):.2e}   greedy tokens {'same' if tok == ref_tok else 'CHANGED'}")
    ow.WhisperOracle._sdpa = orig


if __name__ == "__main__":
    main()
