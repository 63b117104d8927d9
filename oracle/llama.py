"""Oracle for the Orpheus (Llama-3 geometry) TTS path, SURVEY.md section 8 rows a7-a13.
Test infrastructure only.

Follows:
  Sources/MLXAudioTTS/Models/Llama/LlamaTTS.swift:41-98     SNAC 7-token frame (de)interleave
  Sources/MLXAudioTTS/Models/Llama/LlamaTTS.swift:104-202   Llama3ScaledRoPE (freqs used as DIVISOR)
  Sources/MLXAudioTTS/Models/Llama/LlamaTTS.swift:206-346   attention / MLP / block / inner model
  Sources/MLXAudioTTS/Models/Llama/LlamaTTS.swift:383-434   parseOutput
  Sources/MLXAudioTTS/Models/Llama/LlamaTTS.swift:446-553   prepareInputIds (token-id level)
  Sources/MLXAudioTTS/Models/Llama/LlamaTTS.swift:557-567   tied lm head
  Sources/MLXAudioTTS/Models/Llama/LlamaTTS.swift:658-765   generate loop
Sampler / repetition penalty live in the un-vendored dependency mlx-swift-lm 3.31.4
(``MLXLMCommon``: ``GenerateParameters.sampler()/processor()``; call sites LlamaTTS.swift:691-692);
their published algorithm is restated in ``repetition_penalty`` / ``top_p_filter`` below:
  RepetitionContext.process : logits[tok] = l<0 ? l*penalty : l/penalty for tok in last N tokens
  TopPSampler               : p = softmax(l/temp); sort ascending; keep where cumsum > 1-topP;
                              draw categorical over the kept probabilities.

Numerics.  Weights are bf16 (as shipped).  ``round_acts=False`` evaluates the reference's graph with
fp32 activations and fp32 accumulation -- the value the reference's bf16 pipeline approximates; the
device path (bf16 hi/lo activation pairs, fp32 KV cache, fp32 accumulate) is compared with THIS
(1e-3 relative; it actually tracks it to ~1e-5).  ``round_acts=True`` additionally rounds activations
to bf16 at every Linear input, q and the K/V cache (a lower bound on what MLX itself does, which
rounds every op output); it is kept to show the scale of bf16-activation noise (~1e-2 on logits).
Two bf16-activation pipelines cannot be compared at 1e-3: a 1e-7 difference before a rounding point
becomes sqrt(1e-7 * 2^-8) ~ 2e-5 after it and saturates near 1e-3 within two layers (measured, see
DESIGN.md section 2), which is why parity is anchored on the fp32-activation value.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

# LlamaTTS.swift:20-31
START_OF_HUMAN, END_OF_HUMAN, END_OF_TEXT = 128259, 128260, 128009
START_OF_SPEECH, END_OF_SPEECH, PAD_TOKEN = 128257, 128258, 128263
AUDIO_START, AUDIO_END, AUDIO_TOKEN_OFFSET = 128261, 128262, 128266


@dataclass
class LlamaConfig:
    """LlamaTTSConfig.swift:15-60; defaults = Orpheus-3B (Llama-3.2-3B geometry)."""
    hidden_size: int = 3072
    num_hidden_layers: int = 28
    intermediate_size: int = 8192
    num_attention_heads: int = 24
    num_key_value_heads: int = 8
    head_dim: int = 128
    rms_norm_eps: float = 1e-5
    vocab_size: int = 156940
    rope_theta: float = 500000.0
    rope_factor: float = 32.0
    rope_low_freq_factor: float = 1.0
    rope_high_freq_factor: float = 4.0
    rope_old_context_len: float = 8192.0
    tie_word_embeddings: bool = True

    @staticmethod
    def tiny(vocab: int = AUDIO_TOKEN_OFFSET + 7 * 4096 + 2) -> "LlamaConfig":
        return LlamaConfig(hidden_size=256, num_hidden_layers=2, intermediate_size=512,
                           num_attention_heads=4, num_key_value_heads=2, head_dim=64, vocab_size=vocab)


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def init_weights(cfg: LlamaConfig, seed: int = 1234, std: float = 0.02) -> Dict[str, torch.Tensor]:
    """Random init N(0, std^2) in bf16, norm gains 1 +- 0.1 (HF key names, [out,in] row-major)."""
    g = torch.Generator().manual_seed(seed)
    H, I, hd = cfg.hidden_size, cfg.intermediate_size, cfg.head_dim
    nq, nkv = cfg.num_attention_heads, cfg.num_key_value_heads

    def lin(o, i, s=std):
        return (torch.randn(o, i, generator=g) * s).to(torch.bfloat16)

    def gain(n):
        return (1.0 + 0.1 * torch.randn(n, generator=g)).to(torch.bfloat16)

    w = {"model.embed_tokens.weight": lin(cfg.vocab_size, H)}
    for l in range(cfg.num_hidden_layers):
        p = f"model.layers.{l}."
        w[p + "self_attn.q_proj.weight"] = lin(nq * hd, H)
        w[p + "self_attn.k_proj.weight"] = lin(nkv * hd, H)
        w[p + "self_attn.v_proj.weight"] = lin(nkv * hd, H)
        w[p + "self_attn.o_proj.weight"] = lin(H, nq * hd)
        w[p + "mlp.gate_proj.weight"] = lin(I, H)
        w[p + "mlp.up_proj.weight"] = lin(I, H)
        w[p + "mlp.down_proj.weight"] = lin(H, I)
        w[p + "input_layernorm.weight"] = gain(H)
        w[p + "post_attention_layernorm.weight"] = gain(H)
    w["model.norm.weight"] = gain(H)
    if not cfg.tie_word_embeddings:
        w["lm_head.weight"] = lin(cfg.vocab_size, H)
    return w


def llama3_rope_freqs(cfg: LlamaConfig) -> np.ndarray:
    """LlamaTTS.swift:121-156, float32: freqs = base^(i/d) (a DIVISOR of position), long wavelengths
    scaled by ``factor``, medium ones smoothly interpolated."""
    f32 = np.float32
    d = cfg.head_dim
    idx = np.arange(0, d, 2, dtype=f32)
    freqs = np.power(f32(cfg.rope_theta), idx / f32(d), dtype=f32)
    wavelens = (f32(2.0 * np.pi) * freqs).astype(f32)
    low_wl = f32(cfg.rope_old_context_len / cfg.rope_low_freq_factor)
    high_wl = f32(cfg.rope_old_context_len / cfg.rope_high_freq_factor)
    freqs = np.where(wavelens > low_wl, freqs * f32(cfg.rope_factor), freqs).astype(f32)
    is_med = (wavelens > high_wl) & (wavelens < low_wl)
    smooth = ((f32(cfg.rope_old_context_len) / wavelens - f32(cfg.rope_low_freq_factor))
              / f32(cfg.rope_high_freq_factor - cfg.rope_low_freq_factor)).astype(f32)
    denom = ((f32(1.0) - smooth) / f32(cfg.rope_factor) + smooth).astype(f32)
    return np.where(is_med, freqs / denom, freqs).astype(f32)


def rope(x: torch.Tensor, positions: torch.Tensor, freqs: torch.Tensor) -> torch.Tensor:
    """MLXFast.RoPE(traditional:false, freqs:) (LlamaTTS.swift:192-200): angle = pos / freqs[i],
    pairs (i, i + d/2).  x [B, heads, L, d]; positions [L]."""
    d2 = x.shape[-1] // 2
    ang = positions[:, None].to(torch.float32) / freqs[None, :]
    cos, sin = torch.cos(ang), torch.sin(ang)
    x1, x2 = x[..., :d2], x[..., d2:]
    return torch.cat([x1 * cos - x2 * sin, x2 * cos + x1 * sin], dim=-1)


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    return x * torch.rsqrt((x * x).mean(-1, keepdim=True) + eps) * w.to(torch.float32)


class LlamaOracle:
    """Stateful forward with a contiguous KV cache (KVCacheSimple semantics, LlamaTTS.swift:604-608)."""

    def __init__(self, cfg: LlamaConfig, weights: Dict[str, torch.Tensor], round_acts: bool = True):
        self.cfg, self.w, self.round = cfg, weights, round_acts
        self.freqs = torch.from_numpy(llama3_rope_freqs(cfg))
        self.reset()

    def reset(self):
        self.k = [None] * self.cfg.num_hidden_layers
        self.v = [None] * self.cfg.num_hidden_layers
        self.offset = 0

    def _r(self, x):
        return bf16_round(x) if self.round else x

    def _lin(self, x, name):
        return self._r(x) @ self.w[name].to(torch.float32).T

    @torch.no_grad()
    def forward(self, ids: torch.Tensor, trace: Optional[list] = None,
                head_positions: Optional[Sequence[int]] = None) -> torch.Tensor:
        """ids [B, L] -> logits [B, L, V] (LlamaTTS.swift:335-345, 557-567).  ``trace`` (a list) receives the
        residual stream of the LAST position at every RMSNorm input (2 per layer + final), for debugging.
        ``head_positions`` (indices into L) restricts the lm head to those positions -> [B, len, V]; the full-width
        parity tests use it so a 156 940-row head is not evaluated at every teacher-forced position."""
        cfg = self.cfg
        B, L = ids.shape
        nq, nkv, hd = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
        h = self.w["model.embed_tokens.weight"][ids].to(torch.float32)
        pos = torch.arange(self.offset, self.offset + L)
        for l in range(cfg.num_hidden_layers):
            p = f"model.layers.{l}."
            if trace is not None:
                trace.append(h[:, -1].clone())
            xn = rms_norm(h, self.w[p + "input_layernorm.weight"], cfg.rms_norm_eps)
            q = self._lin(xn, p + "self_attn.q_proj.weight").view(B, L, nq, hd).transpose(1, 2)
            k = self._lin(xn, p + "self_attn.k_proj.weight").view(B, L, nkv, hd).transpose(1, 2)
            v = self._lin(xn, p + "self_attn.v_proj.weight").view(B, L, nkv, hd).transpose(1, 2)
            q = self._r(rope(q, pos, self.freqs))
            k = self._r(rope(k, pos, self.freqs))
            v = self._r(v)
            self.k[l] = k if self.k[l] is None else torch.cat([self.k[l], k], dim=2)
            self.v[l] = v if self.v[l] is None else torch.cat([self.v[l], v], dim=2)
            kk = self.k[l].repeat_interleave(nq // nkv, dim=1)
            vv = self.v[l].repeat_interleave(nq // nkv, dim=1)
            s = (q @ kk.transpose(-1, -2)) * (hd ** -0.5)
            S = kk.shape[2]
            if L > 1:   # createAttentionMask: causal w.r.t. absolute positions (LlamaTTS.swift:338)
                mask = torch.arange(S)[None, :] > pos[:, None]
                s = s.masked_fill(mask, float("-inf"))
            a = torch.softmax(s, dim=-1) @ vv
            a = a.transpose(1, 2).reshape(B, L, nq * hd)
            h = h + self._lin(a, p + "self_attn.o_proj.weight")
            if trace is not None:
                trace.append(h[:, -1].clone())
            xn = rms_norm(h, self.w[p + "post_attention_layernorm.weight"], cfg.rms_norm_eps)
            g = self._lin(xn, p + "mlp.gate_proj.weight")
            u = self._lin(xn, p + "mlp.up_proj.weight")
            h = h + self._lin(torch.nn.functional.silu(g) * u, p + "mlp.down_proj.weight")
        self.offset += L
        if trace is not None:
            trace.append(h[:, -1].clone())
        if head_positions is not None:
            h = h[:, list(head_positions)]
        hn = rms_norm(h, self.w["model.norm.weight"], cfg.rms_norm_eps)
        head = self.w["model.embed_tokens.weight"] if cfg.tie_word_embeddings else self.w["lm_head.weight"]
        return self._r(hn) @ head.to(torch.float32).T


# --------------------------------------------------------------------------- sampling (mlx-swift-lm)

def repetition_penalty(logits: np.ndarray, context: Sequence[int], penalty: float) -> np.ndarray:
    """RepetitionContext.process: penalise the (unique) tokens in ``context`` (last N tokens)."""
    out = np.array(logits, dtype=np.float32, copy=True)
    if len(context) and penalty != 1.0:
        idx = np.unique(np.asarray(context, dtype=np.int64))
        sel = out[idx]
        out[idx] = np.where(sel < 0, sel * np.float32(penalty), sel / np.float32(penalty))
    return out


def top_p_filter(logits: np.ndarray, temperature: float, top_p: float) -> np.ndarray:
    """TopPSampler: returns the (unnormalised) kept probabilities, zero elsewhere."""
    l = np.asarray(logits, dtype=np.float64) / temperature
    p = np.exp(l - l.max())
    p /= p.sum()
    order = np.argsort(p, kind="stable")
    cum = np.cumsum(p[order])
    keep_sorted = cum > (1.0 - top_p)
    out = np.zeros_like(p)
    out[order[keep_sorted]] = p[order[keep_sorted]]
    return out


def sample_inverse_cdf(kept: np.ndarray, u: float) -> int:
    """Deterministic categorical draw used by the device sampler: walk the kept probabilities in
    index order, return the first index whose running sum exceeds u * total."""
    c = np.cumsum(kept)
    return int(np.searchsorted(c, u * c[-1], side="right"))


# --------------------------------------------------------------------------- token plumbing

def prepare_input_ids(prompt_token_ids: List[List[int]]) -> Tuple[np.ndarray, np.ndarray]:
    """LlamaTTS.swift:471-553 without the tokenizer / voice-cloning branches: left-pad with 128263 to
    the longest prompt, frame as [SOH] ids [EOT, EOH].  Returns (ids [B, L+3], mask)."""
    max_len = max((len(p) for p in prompt_token_ids), default=0)
    rows = []
    for p in prompt_token_ids:
        rows.append([PAD_TOKEN] * (max_len - len(p)) + [START_OF_HUMAN] + list(p) + [END_OF_TEXT, END_OF_HUMAN])
    ids = np.asarray(rows, dtype=np.int32)
    return ids, ids != PAD_TOKEN


def parse_output(input_ids: np.ndarray) -> List[List[int]]:
    """LlamaTTS.swift:383-434: crop after the LAST 128257 column found anywhere in the batch, drop
    128258, trim each row to a multiple of 7, subtract 128266."""
    ids = np.asarray(input_ids)
    last = None
    for i in range(ids.shape[0]):
        for j in range(ids.shape[1]):
            if ids[i, j] == START_OF_SPEECH:
                last = j
    cropped = ids[:, last + 1:] if last is not None else ids
    out = []
    for row in cropped:
        r = [int(t) for t in row if t != END_OF_SPEECH]
        r = r[:(len(r) // 7) * 7]
        out.append([t - AUDIO_TOKEN_OFFSET for t in r])
    return out


def codes_from_code_list(code_list: Sequence[int]) -> List[np.ndarray]:
    """llamaDecodeAudioFromCodes, LlamaTTS.swift:41-63: 7-token frame -> 3 SNAC code layers [1,T_i]."""
    l1, l2, l3 = [], [], []
    for i in range((len(code_list) + 1) // 7):
        b = 7 * i
        l1.append(code_list[b])
        l2.append(code_list[b + 1] - 4096)
        l3.append(code_list[b + 2] - 2 * 4096)
        l3.append(code_list[b + 3] - 3 * 4096)
        l2.append(code_list[b + 4] - 4 * 4096)
        l3.append(code_list[b + 5] - 5 * 4096)
        l3.append(code_list[b + 6] - 6 * 4096)
    return [np.asarray(x, dtype=np.int32)[None] for x in (l1, l2, l3)]


def code_list_from_codes(codes: List[np.ndarray]) -> List[int]:
    """llamaEncodeAudioToCodes, LlamaTTS.swift:72-98 (inverse interleave)."""
    l1, l2, l3 = (np.asarray(c).reshape(-1) for c in codes)
    out = []
    for i in range(len(l1)):
        out += [int(l1[i]), int(l2[2 * i]) + 4096, int(l3[4 * i]) + 2 * 4096, int(l3[4 * i + 1]) + 3 * 4096,
                int(l2[2 * i + 1]) + 4 * 4096, int(l3[4 * i + 2]) + 5 * 4096, int(l3[4 * i + 3]) + 6 * 4096]
    return out


@torch.no_grad()
def generate_tokens(model: LlamaOracle, input_ids: np.ndarray, max_tokens: int, temperature: float = 0.0,
                    top_p: float = 1.0, rep_penalty: float = 1.0, rep_context: int = 20,
                    uniforms: Optional[np.ndarray] = None, mask_eos: bool = False,
                    return_logits: bool = False):
    """LlamaTTS.swift:683-744 per utterance (the reference is batch-1; a batch here is B independent
    utterances, "batched == serial").  Greedy when temperature == 0 (lowest index wins ties).
    ``uniforms[b, step]`` feed ``sample_inverse_cdf``.  Returns list of generated-token lists."""
    B = input_ids.shape[0]
    model.reset()
    logits = model.forward(torch.as_tensor(input_ids, dtype=torch.long))[:, -1].numpy()
    ctx = [list(map(int, row[-rep_context:])) if rep_context > 0 else [] for row in input_ids]
    done = [False] * B
    gen: List[List[int]] = [[] for _ in range(B)]
    all_logits = []
    for step in range(max_tokens):
        nxt = np.zeros(B, dtype=np.int64)
        proc = np.empty_like(logits, dtype=np.float32)
        for b in range(B):
            l = repetition_penalty(logits[b], ctx[b], rep_penalty)
            if mask_eos:
                l[END_OF_SPEECH] = -np.inf
            proc[b] = l
            if temperature == 0.0:
                t = int(np.argmax(l))
            else:
                kept = top_p_filter(l, temperature, top_p)
                t = sample_inverse_cdf(kept, float(uniforms[b, step]))
            nxt[b] = t
        all_logits.append(proc)
        for b in range(B):
            if done[b]:
                continue
            if nxt[b] == END_OF_SPEECH:
                done[b] = True
            else:
                gen[b].append(int(nxt[b]))
                if rep_context > 0:
                    ctx[b] = (ctx[b] + [int(nxt[b])])[-rep_context:]
        if all(done):
            break
        logits = model.forward(torch.as_tensor(nxt[:, None], dtype=torch.long))[:, -1].numpy()
    return (gen, all_logits) if return_logits else gen
