"""Oracle for the Whisper STT path.  Test infrastructure only.

Follows:
  Sources/MLXAudioSTT/Models/Whisper/WhisperLayers.swift:11-73    WhisperAttention (k_proj has no bias)
  Sources/MLXAudioSTT/Models/Whisper/WhisperLayers.swift:77-156   encoder layer / encoder (conv stem, exact GELU)
  Sources/MLXAudioSTT/Models/Whisper/WhisperLayers.swift:169-328  decoder layer / decoder (self KV concat, cached cross K/V)
  Sources/MLXAudioSTT/Models/Whisper/WhisperModel.swift:186-309   transcribeChunk greedy loop, suppress masks
  Sources/MLXAudioSTT/Models/Whisper/WhisperModel.swift:384-397   whisperSinusoids
  Sources/MLXAudioSTT/Models/Whisper/WhisperTokenizer.swift:98-113 buildPromptTokens (token-id level)
Weights use the HuggingFace key names and layouts (``model.encoder.conv1.weight`` is PyTorch ``[out, in, k]``; the
reference transposes it to MLX ``[out, k, in]`` in ``sanitizeHuggingFace``, WhisperModel.swift:335-365).
Numerics as ``oracle/llama.py``: bf16 weights, fp32 activations and accumulation (``round_acts`` unused here).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from . import dsp

# multilingual vocabulary ids (WhisperTokenizer.swift; openai/whisper-base added_tokens)
EOT, SOT, TRANSLATE, TRANSCRIBE, NO_TIMESTAMPS, TIMESTAMP_BEGIN = 50257, 50258, 50358, 50359, 50363, 50364
LANG_EN = 50259


@dataclass
class WhisperConfig:
    """WhisperConfig.swift:58-76 keys; defaults = whisper-base."""
    vocab_size: int = 51865
    num_mel_bins: int = 80
    d_model: int = 512
    encoder_layers: int = 6
    encoder_attention_heads: int = 8
    encoder_ffn_dim: int = 2048
    max_source_positions: int = 1500
    decoder_layers: int = 6
    decoder_attention_heads: int = 8
    decoder_ffn_dim: int = 2048
    max_target_positions: int = 448

    @staticmethod
    def tiny_test() -> "WhisperConfig":
        return WhisperConfig(vocab_size=51865, d_model=128, encoder_layers=2, encoder_attention_heads=2, encoder_ffn_dim=256,
                             decoder_layers=2, decoder_attention_heads=2, decoder_ffn_dim=256)


def sinusoids(length: int, channels: int) -> torch.Tensor:
    """WhisperModel.swift:384-397."""
    half = channels // 2
    inc = math.log(10000.0) / max(half - 1, 1)
    t = torch.arange(length, dtype=torch.float64)[:, None] * torch.exp(-inc * torch.arange(half, dtype=torch.float64))[None]
    return torch.cat([torch.sin(t), torch.cos(t)], dim=1).to(torch.float32)


def init_weights(cfg: WhisperConfig, seed: int = 1234, std: float = 0.05) -> Dict[str, torch.Tensor]:
    """Random init: matrices N(0, std^2) in bf16; biases / LayerNorm / positions in fp32."""
    g = torch.Generator().manual_seed(seed)
    d = cfg.d_model

    def mat(*shape, s=std):
        return (torch.randn(*shape, generator=g) * s).to(torch.bfloat16)

    def vec(n, s=0.02, base=0.0):
        return (base + s * torch.randn(n, generator=g)).to(torch.float32)

    w: Dict[str, torch.Tensor] = {}
    w["model.encoder.conv1.weight"] = mat(d, cfg.num_mel_bins, 3)
    w["model.encoder.conv1.bias"] = vec(d)
    w["model.encoder.conv2.weight"] = mat(d, d, 3, s=std / 2)
    w["model.encoder.conv2.bias"] = vec(d)
    w["model.encoder.embed_positions.weight"] = sinusoids(cfg.max_source_positions, d)
    w["model.decoder.embed_tokens.weight"] = mat(cfg.vocab_size, d)
    w["model.decoder.embed_positions.weight"] = vec(cfg.max_target_positions * d, 0.02).view(cfg.max_target_positions, d)

    def attn(p):
        w[p + "q_proj.weight"], w[p + "q_proj.bias"] = mat(d, d), vec(d)
        w[p + "k_proj.weight"] = mat(d, d)
        w[p + "v_proj.weight"], w[p + "v_proj.bias"] = mat(d, d), vec(d)
        w[p + "out_proj.weight"], w[p + "out_proj.bias"] = mat(d, d), vec(d)

    def ln(p):
        w[p + "weight"], w[p + "bias"] = vec(d, 0.1, 1.0), vec(d, 0.05)

    for l in range(cfg.encoder_layers):
        p = f"model.encoder.layers.{l}."
        attn(p + "self_attn."); ln(p + "self_attn_layer_norm.")
        w[p + "fc1.weight"], w[p + "fc1.bias"] = mat(cfg.encoder_ffn_dim, d), vec(cfg.encoder_ffn_dim)
        w[p + "fc2.weight"], w[p + "fc2.bias"] = mat(d, cfg.encoder_ffn_dim), vec(d)
        ln(p + "final_layer_norm.")
    ln("model.encoder.layer_norm.")
    for l in range(cfg.decoder_layers):
        p = f"model.decoder.layers.{l}."
        attn(p + "self_attn."); ln(p + "self_attn_layer_norm.")
        attn(p + "encoder_attn."); ln(p + "encoder_attn_layer_norm.")
        w[p + "fc1.weight"], w[p + "fc1.bias"] = mat(cfg.decoder_ffn_dim, d), vec(cfg.decoder_ffn_dim)
        w[p + "fc2.weight"], w[p + "fc2.bias"] = mat(d, cfg.decoder_ffn_dim), vec(d)
        ln(p + "final_layer_norm.")
    ln("model.decoder.layer_norm.")
    return w


class WhisperOracle:
    def __init__(self, cfg: WhisperConfig, weights: Dict[str, torch.Tensor]):
        self.cfg = cfg
        self.w = {k: v.to(torch.float32) for k, v in weights.items()}
        self.reset()

    def reset(self):
        L = self.cfg.decoder_layers
        self.self_k: List[Optional[torch.Tensor]] = [None] * L
        self.self_v: List[Optional[torch.Tensor]] = [None] * L
        self.cross_k: List[Optional[torch.Tensor]] = [None] * L
        self.cross_v: List[Optional[torch.Tensor]] = [None] * L

    def _lin(self, x, p, bias=True):
        y = x @ self.w[p + "weight"].T
        return y + self.w[p + "bias"] if bias else y

    def _ln(self, x, p):
        return F.layer_norm(x, (x.shape[-1],), self.w[p + "weight"], self.w[p + "bias"], 1e-5)

    def _heads(self, x, nh):
        B, T, D = x.shape
        return x.view(B, T, nh, D // nh).transpose(1, 2)

    def _sdpa(self, q, k, v, nh, mask=None):
        hd = q.shape[-1]
        s = (q @ k.transpose(-1, -2)) * hd ** -0.5               # WhisperLayers.swift:21,62-68
        if mask is not None:
            s = s + mask
        o = torch.softmax(s, dim=-1) @ v
        B, _, T, _ = o.shape
        return o.transpose(1, 2).reshape(B, T, nh * hd)

    @torch.no_grad()
    def encode(self, features: torch.Tensor) -> torch.Tensor:
        """features [B, 3000, n_mels] (WhisperAudio.encoderFeatures) -> [B, 1500, d]   (WhisperLayers.swift:146-155)"""
        cfg = self.cfg
        x = features.to(torch.float32).transpose(1, 2)
        h = F.gelu(F.conv1d(x, self.w["model.encoder.conv1.weight"], self.w["model.encoder.conv1.bias"], padding=1))
        h = F.gelu(F.conv1d(h, self.w["model.encoder.conv2.weight"], self.w["model.encoder.conv2.bias"], stride=2, padding=1))
        h = h.transpose(1, 2)
        h = h + self.w["model.encoder.embed_positions.weight"][:h.shape[1]]
        nh = cfg.encoder_attention_heads
        for l in range(cfg.encoder_layers):
            p = f"model.encoder.layers.{l}."
            y = self._ln(h, p + "self_attn_layer_norm.")
            q = self._heads(self._lin(y, p + "self_attn.q_proj."), nh)
            k = self._heads(self._lin(y, p + "self_attn.k_proj.", bias=False), nh)
            v = self._heads(self._lin(y, p + "self_attn.v_proj."), nh)
            h = h + self._lin(self._sdpa(q, k, v, nh), p + "self_attn.out_proj.")
            y = self._ln(h, p + "final_layer_norm.")
            h = h + self._lin(F.gelu(self._lin(y, p + "fc1.")), p + "fc2.")
        return self._ln(h, "model.encoder.layer_norm.")

    @torch.no_grad()
    def decode(self, tokens: torch.Tensor, start: int, enc: torch.Tensor) -> torch.Tensor:
        """tokens [B, Tnew] at positions start.. -> hidden [B, Tnew, d] (WhisperLayers.swift:282-312)."""
        cfg = self.cfg
        nh = cfg.decoder_attention_heads
        Tn = tokens.shape[1]
        h = self.w["model.decoder.embed_tokens.weight"][tokens] + self.w["model.decoder.embed_positions.weight"][start:start + Tn]
        mask = None
        if Tn > 1:
            total = start + Tn
            rows = torch.arange(start, total)[:, None]
            cols = torch.arange(total)[None, :]
            mask = torch.where(cols <= rows, 0.0, -1e9)
        for l in range(cfg.decoder_layers):
            p = f"model.decoder.layers.{l}."
            y = self._ln(h, p + "self_attn_layer_norm.")
            q = self._heads(self._lin(y, p + "self_attn.q_proj."), nh)
            k = self._heads(self._lin(y, p + "self_attn.k_proj.", bias=False), nh)
            v = self._heads(self._lin(y, p + "self_attn.v_proj."), nh)
            self.self_k[l] = k if self.self_k[l] is None else torch.cat([self.self_k[l], k], dim=2)
            self.self_v[l] = v if self.self_v[l] is None else torch.cat([self.self_v[l], v], dim=2)
            h = h + self._lin(self._sdpa(q, self.self_k[l], self.self_v[l], nh, mask), p + "self_attn.out_proj.")
            y = self._ln(h, p + "encoder_attn_layer_norm.")
            if self.cross_k[l] is None:
                self.cross_k[l] = self._heads(self._lin(enc, p + "encoder_attn.k_proj.", bias=False), nh)
                self.cross_v[l] = self._heads(self._lin(enc, p + "encoder_attn.v_proj."), nh)
            q = self._heads(self._lin(y, p + "encoder_attn.q_proj."), nh)
            h = h + self._lin(self._sdpa(q, self.cross_k[l], self.cross_v[l], nh), p + "encoder_attn.out_proj.")
            y = self._ln(h, p + "final_layer_norm.")
            h = h + self._lin(F.gelu(self._lin(y, p + "fc1.")), p + "fc2.")
        return self._ln(h, "model.decoder.layer_norm.")

    def logits(self, hidden: torch.Tensor) -> torch.Tensor:
        return hidden @ self.w["model.decoder.embed_tokens.weight"].T          # tied (WhisperLayers.swift:325)


def build_prompt_tokens(language_id: Optional[int] = LANG_EN, task: str = "transcribe", multilingual: bool = True) -> List[int]:
    """WhisperTokenizer.swift:98-113 with the language already resolved to its token id."""
    toks = [SOT]
    if multilingual:
        if language_id is not None:
            toks.append(language_id)
        toks.append(TRANSLATE if task.lower() == "translate" else TRANSCRIBE)
    toks.append(NO_TIMESTAMPS)
    return toks


@torch.no_grad()
def transcribe_tokens(model: WhisperOracle, audio: np.ndarray, prompt: Sequence[int], max_tokens: int = 432,
                      begin_suppress: Sequence[int] = (EOT,), suppress: Sequence[int] = (), mask_eot: bool = False,
                      return_logits: bool = False):
    """transcribeChunk (WhisperModel.swift:186-282), greedy: one <=30 s clip -> generated token ids."""
    cfg = model.cfg
    feats = torch.from_numpy(dsp.whisper_encoder_features(audio, cfg.num_mel_bins)).to(torch.float32)
    model.reset()
    enc = model.encode(feats)
    ids = torch.as_tensor([list(prompt)], dtype=torch.long)
    logits = model.logits(model.decode(ids, 0, enc)[0, -1])
    max_tokens = max(1, min(max_tokens, cfg.max_target_positions - len(prompt) - 1))
    out, all_logits = [], []
    for step in range(max_tokens):
        l = logits.clone()
        if step == 0:
            for i in begin_suppress:
                l[i] += -1e9
        for i in suppress:
            l[i] += -1e9
        l[TIMESTAMP_BEGIN:] += -1e9                                       # suppressFromIndex (:301-309)
        if mask_eot:
            l[EOT] = -float("inf")
        all_logits.append(l.numpy())
        nxt = int(torch.argmax(l))
        if nxt == EOT:
            break
        out.append(nxt)
        logits = model.logits(model.decode(torch.as_tensor([[nxt]]), len(prompt) + step, enc)[0, -1])
    return (out, all_logits) if return_logits else out
