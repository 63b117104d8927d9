"""Oracle for the Qwen3-TTS talker + code-predictor step (BASELINE config 5).  Test infrastructure only;
there is NO CUDA path for this row yet -- this is the "oracle first" step of it.

Follows (paths relative to the reference checkout):
  Sources/MLXAudioTTS/Models/Qwen3TTS/Qwen3TTSTalker.swift:8-34      rotateHalf / applyRotaryPosEmb / computeInvFreq
  Sources/MLXAudioTTS/Models/Qwen3TTS/Qwen3TTSTalker.swift:44-104    TalkerRotaryEmbedding (interleaved 3-section MRoPE)
  Sources/MLXAudioTTS/Models/Qwen3TTS/Qwen3TTSTalker.swift:127-186   TalkerAttention (per-head q/k RMSNorm BEFORE RoPE, GQA, SDPA)
  Sources/MLXAudioTTS/Models/Qwen3TTS/Qwen3TTSTalker.swift:190-245   TalkerMLP (SwiGLU), ResizeMLP, TalkerDecoderLayer
  Sources/MLXAudioTTS/Models/Qwen3TTS/Qwen3TTSTalker.swift:249-310   Qwen3TTSTalkerModel (inputs are EMBEDDINGS; causal mask for L > 1)
  Sources/MLXAudioTTS/Models/Qwen3TTS/Qwen3TTSTalker.swift:314-366   ...ForConditionalGeneration (codec_head, sanitize "talker." prefix)
  Sources/MLXAudioTTS/Models/Qwen3TTS/Qwen3TTSCodePredictor.swift:14-243  code predictor (standard RoPE, 15 lm heads / embeddings,
                                                                     optional small_to_mtp_projection)
  Sources/MLXAudioTTS/Models/Qwen3TTS/Qwen3TTS.swift:380-495         the frame loop (talker step -> 15 predictor steps with the
                                                                     predictor cache trimmed every frame -> summed-embedding feedback)
  Sources/MLXAudioTTS/Models/Qwen3TTS/Qwen3TTS.swift:1003-1118       sampleToken (suppress, repetition penalty over UNIQUE generated
                                                                     tokens, top-k, top-p, min-p, EOS logit re-inserted, categorical)
  Sources/MLXAudioTTS/Models/Qwen3TTS/Qwen3TTSConfig.swift:45-63,268-292  defaults
Weights use the reference's keys after `sanitize` strips "talker." (`model.layers.N.self_attn.q_norm.weight`, ...,
`code_predictor.model.codec_embedding.N.weight`, `code_predictor.lm_head.N.weight`).  float64 by default.
PARITY UNPINNED against the reference itself (no MLX here); pinned against `transformers` Qwen3Model (the talker backbone with
identical position rows) and Qwen3-VL's apply_interleaved_mrope in tests/test_oracle_qwen3_tts.py.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

DTYPE = torch.float64


@dataclass
class CodePredictorConfig:
    """Qwen3TTSConfig.swift:45-63."""
    vocab_size: int = 2048
    hidden_size: int = 1024
    intermediate_size: int = 3072
    num_hidden_layers: int = 5
    num_attention_heads: int = 16
    num_key_value_heads: int = 8
    head_dim: int = 128
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1_000_000.0
    attention_bias: bool = False
    num_code_groups: int = 16


@dataclass
class TalkerConfig:
    """Qwen3TTSConfig.swift:268-292."""
    vocab_size: int = 3072
    hidden_size: int = 1024
    intermediate_size: int = 3072
    num_hidden_layers: int = 28
    num_attention_heads: int = 16
    num_key_value_heads: int = 8
    head_dim: int = 128
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1_000_000.0
    attention_bias: bool = False
    num_code_groups: int = 16
    text_hidden_size: int = 2048
    text_vocab_size: int = 151936
    codec_eos_token_id: int = 2150
    mrope_section: Sequence[int] = (24, 20, 20)
    code_predictor: CodePredictorConfig = field(default_factory=CodePredictorConfig)


# ------------------------------------------------------------------------------------------------ weights
def init_weights(cfg: TalkerConfig, seed: int = 1234, std: float = 0.02) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)

    def n(*shape):
        return (torch.randn(*shape, generator=g) * std).to(DTYPE)

    def norm(d):
        return (1.0 + 0.1 * torch.randn(d, generator=g)).to(DTYPE)

    W: Dict[str, torch.Tensor] = {}

    def block(prefix, c, hidden):
        nq, nkv, hd = c.num_attention_heads, c.num_key_value_heads, c.head_dim
        W[prefix + "self_attn.q_proj.weight"] = n(nq * hd, hidden)
        W[prefix + "self_attn.k_proj.weight"] = n(nkv * hd, hidden)
        W[prefix + "self_attn.v_proj.weight"] = n(nkv * hd, hidden)
        W[prefix + "self_attn.o_proj.weight"] = n(hidden, nq * hd)
        if c.attention_bias:
            for p, d in (("q", nq * hd), ("k", nkv * hd), ("v", nkv * hd), ("o", hidden)):
                W[prefix + f"self_attn.{p}_proj.bias"] = n(d)
        W[prefix + "self_attn.q_norm.weight"] = norm(hd)
        W[prefix + "self_attn.k_norm.weight"] = norm(hd)
        W[prefix + "mlp.gate_proj.weight"] = n(c.intermediate_size, hidden)
        W[prefix + "mlp.up_proj.weight"] = n(c.intermediate_size, hidden)
        W[prefix + "mlp.down_proj.weight"] = n(hidden, c.intermediate_size)
        W[prefix + "input_layernorm.weight"] = norm(hidden)
        W[prefix + "post_attention_layernorm.weight"] = norm(hidden)

    H = cfg.hidden_size
    W["model.codec_embedding.weight"] = n(cfg.vocab_size, H)
    W["model.text_embedding.weight"] = n(cfg.text_vocab_size, cfg.text_hidden_size)
    for l in range(cfg.num_hidden_layers):
        block(f"model.layers.{l}.", cfg, H)
    W["model.norm.weight"] = norm(H)
    W["text_projection.linear_fc1.weight"] = n(cfg.text_hidden_size, cfg.text_hidden_size)
    W["text_projection.linear_fc1.bias"] = n(cfg.text_hidden_size)
    W["text_projection.linear_fc2.weight"] = n(H, cfg.text_hidden_size)
    W["text_projection.linear_fc2.bias"] = n(H)
    W["codec_head.weight"] = n(cfg.vocab_size, H)
    cp = cfg.code_predictor
    for i in range(cp.num_code_groups - 1):
        W[f"code_predictor.model.codec_embedding.{i}.weight"] = n(cp.vocab_size, H)        # dimensions = TALKER hidden size (:143-145)
        W[f"code_predictor.lm_head.{i}.weight"] = n(cp.vocab_size, cp.hidden_size)
    for l in range(cp.num_hidden_layers):
        block(f"code_predictor.model.layers.{l}.", cp, cp.hidden_size)
    W["code_predictor.model.norm.weight"] = norm(cp.hidden_size)
    if cp.hidden_size != H:
        W["code_predictor.small_to_mtp_projection.weight"] = n(cp.hidden_size, H)
        W["code_predictor.small_to_mtp_projection.bias"] = n(cp.hidden_size)
    return W


def sanitize(weights: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Qwen3TTSTalkerForConditionalGeneration.sanitize (Qwen3TTSTalker.swift:356-365): keep `talker.*`, drop the prefix."""
    return {k[len("talker."):]: v for k, v in weights.items() if k.startswith("talker.")}


# ------------------------------------------------------------------------------------------------ rotary embeddings
def inv_freq(dim: int, base: float) -> torch.Tensor:
    """computeInvFreq (:28-32): 1 / base^(2i/dim) in float32 like the reference, widened afterwards."""
    ar = torch.arange(0, dim, 2, dtype=torch.float32)
    return (1.0 / torch.tensor(base, dtype=torch.float32).pow(ar / dim)).to(DTYPE)


def apply_interleaved_mrope(freqs: torch.Tensor, section: Sequence[int]) -> torch.Tensor:
    """TalkerRotaryEmbedding.applyInterleavedMrope (:58-80): freqs [3, B, T, hd/2] (temporal / height / width position rows);
    index i takes the H row when i % 3 == 1 and i < 3*section[1], the W row when i % 3 == 2 and i < 3*section[2], else the T row."""
    half = freqs.shape[-1]
    idx = torch.arange(half)
    h_mask = (idx % 3 == 1) & (idx < section[1] * 3)
    w_mask = (idx % 3 == 2) & (idx < section[2] * 3)
    out = torch.where(h_mask.view(1, 1, half), freqs[1], freqs[0])
    return torch.where(w_mask.view(1, 1, half), freqs[2], out)


def mrope_cos_sin(position_ids: torch.Tensor, head_dim: int, base: float, section: Sequence[int]) -> Tuple[torch.Tensor, torch.Tensor]:
    """TalkerRotaryEmbedding.callAsFunction (:82-103).  position_ids [B, T] (broadcast to 3 rows) or [3, B, T] -> cos, sin [B, T, hd]."""
    pos = position_ids
    if pos.dim() == 2:
        pos = pos.unsqueeze(0).expand(3, -1, -1)
    freqs = pos.to(DTYPE).unsqueeze(-1) * inv_freq(head_dim, base).view(1, 1, 1, -1)       # [3, B, T, hd/2]
    comb = apply_interleaved_mrope(freqs, section)
    emb = torch.cat([comb, comb], dim=-1)
    return emb.cos(), emb.sin()


def rope_cos_sin(position_ids: torch.Tensor, head_dim: int, base: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """Qwen3TTSRotaryEmbedding (:108-123), code predictor: standard RoPE, cos / sin [B, T, hd]."""
    freqs = position_ids.to(DTYPE).unsqueeze(-1) * inv_freq(head_dim, base).view(1, 1, -1)
    emb = torch.cat([freqs, freqs], dim=-1)
    return emb.cos(), emb.sin()


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], dim=-1)


# ------------------------------------------------------------------------------------------------ transformer
def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * w


class KVCache:
    """KVCacheSimple (mlx-swift-lm): concatenating cache with an `offset`; `trim(n)` drops the n newest positions."""

    def __init__(self):
        self.k: Optional[torch.Tensor] = None
        self.v: Optional[torch.Tensor] = None

    @property
    def offset(self) -> int:
        return 0 if self.k is None else self.k.shape[2]

    def update(self, k, v):
        self.k = k if self.k is None else torch.cat([self.k, k], dim=2)
        self.v = v if self.v is None else torch.cat([self.v, v], dim=2)
        return self.k, self.v

    def trim(self, n: int) -> int:
        n = min(n, self.offset)
        if n > 0:
            keep = self.offset - n
            self.k, self.v = (None, None) if keep == 0 else (self.k[:, :, :keep], self.v[:, :, :keep])
        return n


def _lin(W, name, x):
    y = x @ W[name + ".weight"].to(DTYPE).T
    b = W.get(name + ".bias")
    return y if b is None else y + b.to(DTYPE)


def attention(W, prefix, c, x, cos, sin, mask, cache: Optional[KVCache]):
    """TalkerAttention / CodePredictorAttention: q_norm / k_norm per head, THEN rotate-half RoPE, cache, SDPA scale hd^-1/2."""
    B, T, _ = x.shape
    nq, nkv, hd = c.num_attention_heads, c.num_key_value_heads, c.head_dim
    q = _lin(W, prefix + "q_proj", x).view(B, T, nq, hd)
    k = _lin(W, prefix + "k_proj", x).view(B, T, nkv, hd)
    v = _lin(W, prefix + "v_proj", x).view(B, T, nkv, hd)
    q = rms_norm(q, W[prefix + "q_norm.weight"].to(DTYPE), c.rms_norm_eps).transpose(1, 2)
    k = rms_norm(k, W[prefix + "k_norm.weight"].to(DTYPE), c.rms_norm_eps).transpose(1, 2)
    v = v.transpose(1, 2)
    ce, se = cos.unsqueeze(1), sin.unsqueeze(1)
    q = q * ce + rotate_half(q) * se
    k = k * ce + rotate_half(k) * se
    if cache is not None:
        k, v = cache.update(k, v)
    g = nq // nkv
    kk, vv = k.repeat_interleave(g, dim=1), v.repeat_interleave(g, dim=1)
    s = (q @ kk.transpose(-1, -2)) * hd ** -0.5
    if mask is not None:
        s = s + mask
    o = torch.softmax(s, dim=-1) @ vv
    return _lin(W, prefix + "o_proj", o.transpose(1, 2).reshape(B, T, nq * hd))


def decoder_layer(W, prefix, c, x, cos, sin, mask, cache):
    h = x + attention(W, prefix + "self_attn.", c, rms_norm(x, W[prefix + "input_layernorm.weight"].to(DTYPE), c.rms_norm_eps), cos, sin, mask, cache)
    y = rms_norm(h, W[prefix + "post_attention_layernorm.weight"].to(DTYPE), c.rms_norm_eps)
    gate, up = _lin(W, prefix + "mlp.gate_proj", y), _lin(W, prefix + "mlp.up_proj", y)
    return h + _lin(W, prefix + "mlp.down_proj", torch.nn.functional.silu(gate) * up)


def causal_mask(T: int, offset: int) -> Optional[torch.Tensor]:
    """createAdditiveCausalMask(seqLen) only when seqLen > 1 (:291-294).  With a non-empty cache the reference's [T, T] mask would
    not even broadcast; its only multi-token call is the prefill (offset 0), which is what this restates (keys before the block open)."""
    if T <= 1:
        return None
    m = torch.full((T, T), float("-inf"), dtype=DTYPE).triu(1)
    return torch.cat([torch.zeros((T, offset), dtype=DTYPE), m], dim=1) if offset else m


class Talker:
    def __init__(self, cfg: TalkerConfig, W: Dict[str, torch.Tensor]):
        self.cfg, self.W = cfg, W

    def make_cache(self) -> List[KVCache]:
        return [KVCache() for _ in range(self.cfg.num_hidden_layers)]

    def embed_codec(self, ids: torch.Tensor) -> torch.Tensor:
        return self.W["model.codec_embedding.weight"].to(DTYPE)[ids]

    def embed_text(self, ids: torch.Tensor) -> torch.Tensor:
        """text_embedding -> text_projection (ResizeMLP: fc2(silu(fc1(x))), :209-221)."""
        e = self.W["model.text_embedding.weight"].to(DTYPE)[ids]
        return _lin(self.W, "text_projection.linear_fc2", torch.nn.functional.silu(_lin(self.W, "text_projection.linear_fc1", e)))

    def model(self, inputs_embeds: torch.Tensor, cache: Optional[List[KVCache]] = None, position_ids: Optional[torch.Tensor] = None):
        """Qwen3TTSTalkerModel.callAsFunction (:271-305) -> final-norm hidden states [B, T, H]."""
        c = self.cfg
        x = inputs_embeds.to(DTYPE)
        B, T, _ = x.shape
        off = cache[0].offset if cache else 0
        if position_ids is None:
            position_ids = torch.arange(off, off + T).view(1, T).expand(B, T)
        cos, sin = mrope_cos_sin(position_ids, c.head_dim, c.rope_theta, c.mrope_section)
        mask = causal_mask(T, off)
        for l in range(c.num_hidden_layers):
            x = decoder_layer(self.W, f"model.layers.{l}.", c, x, cos, sin, mask, cache[l] if cache else None)
        return rms_norm(x, self.W["model.norm.weight"].to(DTYPE), c.rms_norm_eps)

    def __call__(self, inputs_embeds, cache=None, position_ids=None):
        """-> (codec logits [B, T, vocab], hidden [B, T, H])  (:340-350)."""
        h = self.model(inputs_embeds, cache, position_ids)
        return _lin(self.W, "codec_head", h), h


class CodePredictor:
    def __init__(self, cfg: TalkerConfig, W: Dict[str, torch.Tensor]):
        self.cfg, self.cp, self.W = cfg, cfg.code_predictor, W

    def make_cache(self) -> List[KVCache]:
        return [KVCache() for _ in range(self.cp.num_hidden_layers)]

    def embed(self, group: int, ids: torch.Tensor) -> torch.Tensor:
        return self.W[f"code_predictor.model.codec_embedding.{group}.weight"].to(DTYPE)[ids]

    def __call__(self, inputs_embeds: torch.Tensor, cache: Optional[List[KVCache]], generation_step: int) -> torch.Tensor:
        """Qwen3TTSCodePredictor.callAsFunction (:222-238): optional projection -> model -> lm_head[generation_step]."""
        cp = self.cp
        x = inputs_embeds.to(DTYPE)
        if "code_predictor.small_to_mtp_projection.weight" in self.W:
            x = _lin(self.W, "code_predictor.small_to_mtp_projection", x)
        B, T, _ = x.shape
        off = cache[0].offset if cache else 0
        pos = torch.arange(off, off + T).view(1, T).expand(B, T)
        cos, sin = rope_cos_sin(pos, cp.head_dim, cp.rope_theta)
        mask = causal_mask(T, off)
        for l in range(cp.num_hidden_layers):
            x = decoder_layer(self.W, f"code_predictor.model.layers.{l}.", cp, x, cos, sin, mask, cache[l] if cache else None)
        x = rms_norm(x, self.W["code_predictor.model.norm.weight"].to(DTYPE), cp.rms_norm_eps)
        return _lin(self.W, f"code_predictor.lm_head.{generation_step}", x)


# ------------------------------------------------------------------------------------------------ prompt embeddings
def prepare_generation_inputs(cfg: TalkerConfig, W, chat_ids: Sequence[int], tts_bos: int, tts_eos: int, tts_pad: int,
                              language_id: Optional[int] = None, speaker_id: Optional[int] = None,
                              instruct_ids: Optional[Sequence[int]] = None, codec_think_id: int = 2154, codec_nothink_id: int = 2155,
                              codec_think_bos_id: int = 2156, codec_think_eos_id: int = 2157, codec_pad_id: int = 2148,
                              codec_bos_id: int = 2149):
    """prepareGenerationInputs (Qwen3TTS.swift:883-999) from token ids (tokenisation stays with the host tokenizer).
    chat_ids = tokens of "<|im_start|>assistant\n{text}<|im_end|>\n<|im_start|>assistant\n"; instruct_ids = tokens of
    "<|im_start|>user\n{instruct}<|im_end|>\n" (VoiceDesign) or None.  Returns (input_embeds [1, L, H], trailing_text_hidden
    [1, n, H], tts_pad_embed [1, 1, H])."""
    t = Talker(cfg, W)
    text = t.embed_text(torch.as_tensor([list(chat_ids)]))                                   # text_projection(text_embedding(ids)) (:898)
    tts = t.embed_text(torch.as_tensor([[tts_bos, tts_eos, tts_pad]]))
    bos_e, eos_e, pad_e = tts[:, 0:1], tts[:, 1:2], tts[:, 2:3]
    if language_id is not None:                                                              # codec prefix (:938-951)
        prefill = [codec_think_id, codec_think_bos_id, language_id, codec_think_eos_id]
    else:
        prefill = [codec_nothink_id, codec_think_bos_id, codec_think_eos_id]
    codec = t.embed_codec(torch.as_tensor([prefill]))
    suffix = t.embed_codec(torch.as_tensor([[codec_pad_id, codec_bos_id]]))
    parts = [codec] + ([t.embed_codec(torch.as_tensor([[speaker_id]]))] if speaker_id is not None else []) + [suffix]
    codec = torch.cat(parts, dim=1)                                                          # (:957-962)
    role = text[:, :3]                                                                       # "<|im_start|>assistant\n"
    pad_count = codec.shape[1] - 2
    combined = torch.cat([pad_e.expand(1, pad_count, -1), bos_e], dim=1) + codec[:, :-1]     # (:976-979)
    pieces = ([t.embed_text(torch.as_tensor([list(instruct_ids)]))] if instruct_ids else []) + [role, combined]
    first_text = text[:, 3:4] + codec[:, -1:]                                                # (:989)
    inputs = torch.cat(pieces + [first_text], dim=1)
    trailing = torch.cat([text[:, 4:text.shape[1] - 5], eos_e], dim=1)                       # tokens 4 .. -5, then tts EOS (:993-996)
    return inputs, trailing, pad_e


# ------------------------------------------------------------------------------------------------ sampling
def filter_logits(logits: torch.Tensor, temperature: float = 0.9, top_p: float = 1.0, top_k: int = 50, repetition_penalty: float = 1.0,
                  generated_tokens: Optional[Sequence[int]] = None, suppress_tokens: Optional[Sequence[int]] = None,
                  eos_token_id: Optional[int] = None, min_p: float = 0.0) -> torch.Tensor:
    """Everything sampleToken (Qwen3TTS.swift:1003-1118) does to the last position's logits [B, V] before `categorical`
    (or before argmax when temperature <= 0, in which case the filters after the repetition penalty are skipped)."""
    x = logits.clone().to(DTYPE)
    V = x.shape[-1]
    if suppress_tokens:
        x[:, list(suppress_tokens)] = float("-inf")
    if generated_tokens and repetition_penalty != 1.0:
        uniq = sorted({t for t in generated_tokens if t < V})
        if uniq:
            sel = x[:, uniq]
            x[:, uniq] = torch.where(sel < 0, sel * repetition_penalty, sel / repetition_penalty)
    if temperature <= 0:
        return x
    eos = x[:, eos_token_id:eos_token_id + 1].clone() if eos_token_id is not None and 0 <= eos_token_id < V else None
    f = x.clone()
    if 0 < top_k < V:
        # mask everything outside the k largest (argPartition of -logits: ties at the boundary are implementation-defined)
        kth = torch.topk(x, top_k, dim=-1).indices
        keep = torch.zeros_like(x, dtype=torch.bool).scatter_(1, kth, True)
        f = torch.where(keep, f, torch.full_like(f, float("-inf")))
    if 0 < top_p < 1.0:
        probs = torch.softmax(f, dim=-1)
        order = torch.argsort(f, dim=-1)                       # ascending
        cum = torch.cumsum(torch.gather(probs, 1, order), dim=-1)
        cum_orig = torch.zeros_like(cum).scatter_(1, order, cum)
        f = torch.where(cum_orig > 1.0 - top_p, f, torch.full_like(f, float("-inf")))
    if min_p > 0.0:
        top = f.max(dim=-1, keepdim=True).values
        f = torch.where(f < top + float(np.log(min_p)), torch.full_like(f, float("-inf")), f)
    if eos is not None:
        f[:, eos_token_id:eos_token_id + 1] = eos
    return f


def sample_token(logits_last: torch.Tensor, generator: Optional[torch.Generator] = None, **kw) -> torch.Tensor:
    """-> [B, 1] token ids.  temperature <= 0: argmax (lowest index wins ties, like MLX argMax)."""
    t = kw.get("temperature", 0.9)
    f = filter_logits(logits_last, **kw)
    if t <= 0:
        return f.argmax(dim=-1, keepdim=True)
    p = torch.softmax(f / t, dim=-1)
    return torch.multinomial(p, 1, generator=generator)


# ------------------------------------------------------------------------------------------------ the frame loop
def generate_codes(cfg: TalkerConfig, W, input_embeds: torch.Tensor, trailing_text_hidden: torch.Tensor, tts_pad_embed: torch.Tensor,
                   max_tokens: int, temperature: float = 0.0, top_p: float = 1.0, top_k: int = 50, repetition_penalty: float = 1.0,
                   min_p: float = 0.0, generator: Optional[torch.Generator] = None, stop_on_eos: bool = True) -> torch.Tensor:
    """Qwen3TTS.swift:380-495 for one utterance.  input_embeds [1, L, H] (the prepared prompt embeddings), trailing_text_hidden
    [1, n, H] (text embeddings still to be fed, one per frame), tts_pad_embed [1, 1, H].  Returns codes [frames, num_code_groups]."""
    talker, pred = Talker(cfg, W), CodePredictor(cfg, W)
    cache, code_cache = talker.make_cache(), pred.make_cache()
    eos = cfg.codec_eos_token_id
    suppress = [t for t in range(max(cfg.vocab_size - 1024, 0), cfg.vocab_size) if t != eos]      # the special-token block (:383-385)
    generated: List[int] = []
    frames: List[torch.Tensor] = []
    trailing_idx = 0
    x = input_embeds.to(DTYPE)
    skw = dict(temperature=temperature, top_p=top_p, top_k=top_k, min_p=min_p)
    for _ in range(max_tokens):
        logits, hidden = talker(x, cache)
        nxt = sample_token(logits[:, -1], generator, repetition_penalty=repetition_penalty, generated_tokens=generated,
                           suppress_tokens=suppress, eos_token_id=eos, **skw)
        codes = [nxt]
        code_hidden = hidden[:, -1:, :]
        for lc in code_cache:
            lc.trim(lc.offset)                                   # a fresh predictor context every frame (:436-438)
        for ci in range(cfg.num_code_groups - 1):
            if ci == 0:
                inp = torch.cat([code_hidden, talker.embed_codec(nxt)], dim=1)
            else:
                inp = pred.embed(ci - 1, codes[-1])
            cl = pred(inp, code_cache, ci)
            codes.append(sample_token(cl[:, -1], generator, **skw))
        text = trailing_text_hidden[:, trailing_idx:trailing_idx + 1] if trailing_idx < trailing_text_hidden.shape[1] else tts_pad_embed
        if trailing_idx < trailing_text_hidden.shape[1]:
            trailing_idx += 1
        emb = talker.embed_codec(nxt)
        for i, code in enumerate(codes[1:]):
            emb = emb + pred.embed(i, code)
        x = text.to(DTYPE) + emb
        tok = int(nxt[0, 0])
        if stop_on_eos and tok == eos:
            break
        generated.append(tok)
        frames.append(torch.cat(codes, dim=1))
    return torch.cat(frames, dim=0) if frames else torch.zeros((0, cfg.num_code_groups), dtype=torch.long)
