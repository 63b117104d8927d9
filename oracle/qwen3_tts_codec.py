"""Oracle for the Qwen3-TTS speech-tokenizer DECODER (codes -> 24 kHz waveform).
Test infrastructure only -- there is no CUDA path for this row yet; the restatement exists so the next round can build to it.

Follows (paths relative to the reference checkout, file = Sources/MLXAudioTTS/Models/Qwen3TTS/Qwen3TTSSpeechTokenizer.swift):
  :9-121      VectorQuantization / ResidualVectorQuantization / ResidualVectorQuantizer / SplitResidualVectorQuantizer .decode
  (Sources/MLXAudioCodecs/Mimi/Quantization.swift:24-33,49-53  EuclideanCodebook: embedding = embedding_sum / max(usage, 1e-5))
  :135-232    CausalConv1d (left pad k_eff - stride, right pad to a whole frame; depthwise variant; streaming ``step``)
  :236-253    SnakeBeta  (x + sin^2(x e^alpha) / (e^beta + 1e-9))
  :257-297    ConvNeXtBlock (causal depthwise k7 -> LayerNorm 1e-6 -> Linear -> exact GELU -> Linear -> gamma -> residual)
  :301-491    DecoderRMSNorm, LayerScale, rotary (rotate-half), DecoderAttention, DecoderMLP, DecoderTransformer
              (input_proj / output_proj around the layers; plain causal mask -- ``sliding_window`` is in the config but
              is not applied by the reference code, so it is not applied here)
  :495-530    DecoderResidualUnit (SnakeBeta, k7 dilated causal conv, SnakeBeta, k1 conv, residual)
  :533-582    DecoderBlockUpsample (transposed conv k = 2r, stride r, trim r on the right; streaming overflow carry)
  :584-638    DecoderBlock (SnakeBeta -> upsample -> residual units with dilation 1, 3, 9)
  :641-731    DecoderInitialConv / DecoderOutputSnake / DecoderOutputConv (causal k7)
  :735-791    CausalTransposeConv1d, UpsampleLayer (transposed conv k = stride = factor, then ConvNeXt)
  :888-1025   Qwen3TTSSpeechTokenizerDecoder (callAsFunction, streamingStep, chunkedDecode)
  :1059-1092  Qwen3TTSSpeechTokenizer.decode / streamingDecode (layout [B, T, Q] -> [B, Q, T], valid lengths)
  :1094-1440  sanitize (decoder keys only: conv transposes, ``upsample.X.Y`` -> ``upsample.X.layers.Y``, codebook stats)
  :1445-1455  checkArrayShapeQwen3
Config defaults: Qwen3TTSConfig.swift:358-385.

Weights use the reference's post-sanitize parameter keys (below the ``decoder.`` module) and MLX layouts: Conv1d and
ConvTransposed1d ``[out, k, in]``, depthwise ``[C, k, 1]``, Linear ``[out, in]``.  float64 signal path.

REFERENCE QUIRK (restated, not repaired).  DecoderBlockUpsample.step (:548-573) runs the transposed conv WITH its bias on
every chunk and adds the previous chunk's carried tail (which already contains the bias) to the new chunk's head, so the
``k - stride`` samples after every chunk boundary receive the bias twice; the one-shot path adds it once.  streamingStep is
therefore equal to the full decode only when those biases are zero, and the audio the reference ships (decodeChunk ->
streamingDecode, Qwen3TTS.swift:214-231; streamingStep, Qwen3TTS.swift:492-500,533-541) depends on the chunking.  A CUDA
path for this row has to reproduce the chunk boundaries it is given; tests pin both facts.

Pinning: no numeric vectors exist in the reference for this model; the building blocks are checked in
tests/test_oracle_qwen3_tts_codec.py against the identically structured ``transformers`` Qwen3-Omni Code2Wav modules
(CausalConvNet, ConvNeXtBlock, SnakeBeta, DecoderResidualUnit, TransformerLayer), against torch conv primitives, and
through the properties the reference's code implies (streamingStep == full decode up to the quirk above, chunkedDecode ==
full decode for a single chunk, causality).
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

DT = torch.float64


@dataclass
class TokenizerDecoderConfig:
    """Qwen3TTSConfig.swift:358-385 defaults."""
    attention_bias: bool = False
    latent_dim: int = 1024
    codebook_dim: int = 512
    codebook_size: int = 2048
    decoder_dim: int = 1536
    hidden_size: int = 512
    intermediate_size: int = 1024
    layer_scale_initial_scale: float = 0.01
    max_position_embeddings: int = 8000
    head_dim: int = 64
    num_attention_heads: int = 16
    num_hidden_layers: int = 8
    num_key_value_heads: int = 16
    num_quantizers: int = 16
    num_semantic_quantizers: int = 1
    rms_norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    sliding_window: int = 72
    upsample_rates: List[int] = field(default_factory=lambda: [8, 5, 4, 3])
    upsampling_ratios: List[int] = field(default_factory=lambda: [2, 2])

    @property
    def total_upsample(self) -> int:             # :902
        return int(np.prod(list(self.upsample_rates) + list(self.upsampling_ratios)))

    @property
    def output_dim(self) -> int:                 # :918
        return self.decoder_dim // (1 << len(self.upsample_rates))


def tiny_config(**kw) -> TokenizerDecoderConfig:
    """A geometry small enough for CPU tests with every structural feature of the default one."""
    base = dict(latent_dim=32, codebook_dim=16, codebook_size=24, decoder_dim=64, hidden_size=24, intermediate_size=40, head_dim=8,
                num_attention_heads=4, num_key_value_heads=2, num_hidden_layers=2, num_quantizers=4, num_semantic_quantizers=1,
                upsample_rates=[4, 3, 2, 2], upsampling_ratios=[2, 2], layer_scale_initial_scale=0.3)
    base.update(kw)
    return TokenizerDecoderConfig(**base)


def mid_config(**kw) -> TokenizerDecoderConfig:
    """The geometry of the (gated) GPU tests and of tests/golden/qwen3_codec.npz: every structural feature of the shipped model,
    every channel count >= 64 (one 64-channel TMA box never exceeds a tensor) with 96 as the non-multiple-of-64 case, and small
    enough for the float64 oracle to decode ~70 code frames in seconds."""
    base = dict(latent_dim=128, codebook_dim=128, codebook_size=64, decoder_dim=768, hidden_size=64, intermediate_size=128, head_dim=32,
                num_attention_heads=4, num_key_value_heads=2, num_hidden_layers=2, num_quantizers=4, num_semantic_quantizers=1,
                upsample_rates=[4, 3, 2], upsampling_ratios=[2, 2], layer_scale_initial_scale=0.3)
    base.update(kw)
    return TokenizerDecoderConfig(**base)


# ---------------------------------------------------------------- weights

def init_weights(cfg: TokenizerDecoderConfig, seed: int = 1234, std: float = 0.08, out_gain: float = 0.02) -> Dict[str, torch.Tensor]:
    """Random weights with the reference's key set and MLX layouts (float32 values, like a checkpoint).  ``out_gain``
    scales the last conv so that the waveform mostly stays inside the final clip(-1, 1)."""
    g = torch.Generator().manual_seed(seed)
    W: Dict[str, torch.Tensor] = {}

    def rn(*shape, s=std):
        return (torch.randn(*shape, generator=g) * s).float()

    def conv(prefix, cout, k, cin, bias=True):
        W[prefix + ".weight"] = rn(cout, k, cin, s=1.0 / math.sqrt(k * cin))
        if bias:
            W[prefix + ".bias"] = rn(cout, s=0.05)

    def lin(prefix, cout, cin, bias=True):
        W[prefix + ".weight"] = rn(cout, cin, s=1.0 / math.sqrt(cin))
        if bias:
            W[prefix + ".bias"] = rn(cout, s=0.05)

    def snake(prefix, c):
        W[prefix + ".alpha"] = rn(c, s=0.3)
        W[prefix + ".beta"] = rn(c, s=0.3)

    half = cfg.codebook_dim // 2
    for name, n in (("rvq_first", cfg.num_semantic_quantizers), ("rvq_rest", cfg.num_quantizers - cfg.num_semantic_quantizers)):
        for i in range(n):
            p = f"quantizer.{name}.vq.layers.{i}.codebook"
            usage = torch.rand(cfg.codebook_size, generator=g).float() * 4.0
            usage[0] = 0.0                                       # exercises the max(usage, eps) clamp
            W[p + ".cluster_usage"] = usage
            W[p + ".embedding_sum"] = rn(cfg.codebook_size, half, s=1.0) * usage.clamp(min=1e-5)[:, None]
        conv(f"quantizer.{name}.output_proj", cfg.codebook_dim, 1, half, bias=False)
    conv("pre_conv.conv", cfg.latent_dim, 3, cfg.codebook_dim)

    H, hd = cfg.hidden_size, cfg.head_dim
    lin("pre_transformer.input_proj", H, cfg.latent_dim)
    lin("pre_transformer.output_proj", cfg.latent_dim, H)
    W["pre_transformer.norm.weight"] = 1.0 + rn(H, s=0.1)
    for i in range(cfg.num_hidden_layers):
        p = f"pre_transformer.layers.{i}"
        lin(p + ".self_attn.q_proj", cfg.num_attention_heads * hd, H, bias=cfg.attention_bias)
        lin(p + ".self_attn.k_proj", cfg.num_key_value_heads * hd, H, bias=cfg.attention_bias)
        lin(p + ".self_attn.v_proj", cfg.num_key_value_heads * hd, H, bias=cfg.attention_bias)
        lin(p + ".self_attn.o_proj", H, cfg.num_attention_heads * hd, bias=cfg.attention_bias)
        lin(p + ".mlp.gate_proj", cfg.intermediate_size, H, bias=False)
        lin(p + ".mlp.up_proj", cfg.intermediate_size, H, bias=False)
        lin(p + ".mlp.down_proj", H, cfg.intermediate_size, bias=False)
        W[p + ".input_layernorm.weight"] = 1.0 + rn(H, s=0.1)
        W[p + ".post_attention_layernorm.weight"] = 1.0 + rn(H, s=0.1)
        W[p + ".self_attn_layer_scale.scale"] = torch.full((H,), cfg.layer_scale_initial_scale) + rn(H, s=0.01)
        W[p + ".mlp_layer_scale.scale"] = torch.full((H,), cfg.layer_scale_initial_scale) + rn(H, s=0.01)

    L = cfg.latent_dim
    for i, f in enumerate(cfg.upsampling_ratios):
        conv(f"upsample.{i}.layers.0.conv", L, f, L)
        p = f"upsample.{i}.layers.1"
        conv(p + ".dwconv.conv", L, 7, 1)
        W[p + ".norm.weight"] = 1.0 + rn(L, s=0.1)
        W[p + ".norm.bias"] = rn(L, s=0.05)
        lin(p + ".pwconv1", 4 * L, L)
        lin(p + ".pwconv2", L, 4 * L)
        W[p + ".gamma"] = rn(L, s=0.3)

    conv("decoder.0.conv", cfg.decoder_dim, 7, L)
    for b, r in enumerate(cfg.upsample_rates):
        cin, cout = cfg.decoder_dim >> b, cfg.decoder_dim >> (b + 1)
        p = f"decoder.{1 + b}.block"
        snake(p + ".0", cin)
        conv(p + ".1.conv", cout, 2 * r, cin)
        for j in (2, 3, 4):
            snake(f"{p}.{j}.act1", cout)
            conv(f"{p}.{j}.conv1.conv", cout, 7, cout)
            snake(f"{p}.{j}.act2", cout)
            conv(f"{p}.{j}.conv2.conv", cout, 1, cout)
    n = len(cfg.upsample_rates)
    snake(f"decoder.{n + 1}", cfg.output_dim)
    conv(f"decoder.{n + 2}.conv", 1, 7, cfg.output_dim)
    W[f"decoder.{n + 2}.conv.weight"] *= out_gain
    return W


def check_array_shape(shape: Tuple[int, ...]) -> bool:
    """checkArrayShapeQwen3 (:1445-1455): True when a 3-d conv weight already looks like the MLX ``[out, k, in]`` layout."""
    if len(shape) != 3:
        return False
    _, d2, d3 = shape
    if d2 == 1:
        return d3 > 64
    if d3 == 1:
        return d2 <= 64
    return d2 < d3


_PREFIXES = ("speech_tokenizer.", "encoder_model.", "decoder_model.")


def sanitize(weights: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Qwen3TTSSpeechTokenizer.sanitize (:1094-1440) restricted to the decoder keys (encoder.* and speaker-encoder keys are
    skipped: the encoder is only used for voice cloning, outside row N1).  Input: a PyTorch-layout checkpoint."""
    out: Dict[str, torch.Tensor] = {}
    books: Dict[str, Dict[str, torch.Tensor]] = {}
    for raw, v in weights.items():
        k = raw
        stripped = True
        while stripped:                                           # stripKnownPrefixes :1118-1133
            stripped = False
            for p in _PREFIXES:
                if k.startswith(p):
                    k, stripped = k[len(p):], True
                    break
        parts = k.split(".")                                      # stripSpeakerEncoderPrefix != nil (Qwen3TTSSpeakerEncoder.swift:345-354)
        if k in ("", "encoder_model", "decoder_model", "speech_tokenizer") or ("speaker_encoder" in parts and parts.index("speaker_encoder") + 1 < len(parts)):
            continue
        if "_codebook.cluster_usage" in k or "_codebook.embedding_sum" in k:     # :1219-1229
            base = k[: k.rfind("._codebook.")]
            books.setdefault(base, {})["cluster_usage" if "cluster_usage" in k else "embedding_sum"] = v
            continue
        if "_codebook.initialized" in k or ".codebook.initialized" in k:
            continue
        if k.startswith("encoder."):
            continue
        is_tconv = ("upsample" in k and ".0.conv.weight" in k) or ("decoder.decoder" in k and "block.1.conv.weight" in k)
        if is_tconv and v.ndim == 3:                              # torch [in, out, k] -> [out, k, in]   :1388-1392
            if not check_array_shape(tuple(v.shape)):
                v = v.permute(1, 2, 0)
        elif "conv.weight" in k and v.ndim == 3:                  # torch [out, in, k] -> [out, k, in]   :1393-1396
            if not check_array_shape(tuple(v.shape)):
                v = v.permute(0, 2, 1)
        elif "_proj.weight" in k and v.ndim == 3:                 # :1397-1401
            if not check_array_shape(tuple(v.shape)):
                v = v.permute(0, 2, 1)
        if "upsample." in k:                                      # :1406-1413
            k = re.sub(r"upsample\.(\d+)\.(\d+)", r"upsample.\1.layers.\2", k)
        out[k] = v.contiguous()
    for base, d in books.items():                                  # :1431-1438
        if "cluster_usage" in d and "embedding_sum" in d:
            out[base + ".codebook.initialized"] = torch.zeros(1)
            out[base + ".codebook.cluster_usage"] = d["cluster_usage"]
            out[base + ".codebook.embedding_sum"] = d["embedding_sum"]
    return out


def strip_decoder_prefix(weights: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """The sanitized keys live under the ``decoder`` child of Qwen3TTSSpeechTokenizer (:1032); drop that first component."""
    return {k[len("decoder."):]: v for k, v in weights.items() if k.startswith("decoder.")}


# ---------------------------------------------------------------- primitives (NCL tensors, float64)

def _w(W, key) -> torch.Tensor:
    return W[key].to(DT)


def conv1d_mlx(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], stride: int = 1, dilation: int = 1) -> torch.Tensor:
    """MLXNN.Conv1d with padding 0 on an NCL tensor; ``w`` is ``[out, k, in]``."""
    return F.conv1d(x, w.permute(0, 2, 1), b, stride=stride, dilation=dilation)


def conv_transpose1d_mlx(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], stride: int) -> torch.Tensor:
    """MLXNN.ConvTransposed1d, padding 0: y[o, t*stride + j] += x[i, t] w[o, j, i]; ``w`` is ``[out, k, in]``."""
    B, cin, T = x.shape
    cout, k, _ = w.shape
    y = torch.zeros(B, cout, (T - 1) * stride + k, dtype=x.dtype)
    for j in range(k):
        y[:, :, j: j + (T - 1) * stride + 1: stride] += torch.einsum("bit,oi->bot", x, w[:, j, :])
    if b is not None:
        y = y + b[None, :, None]
    return y


def extra_padding(length: int, k_eff: int, stride: int) -> int:
    """CausalConv1d.getExtraPadding (:171-175); the frame count is computed in float32 like the reference."""
    pad = k_eff - stride
    n_frames = np.float32(length - k_eff + pad) / np.float32(stride) + np.float32(1)
    ideal = (int(math.ceil(float(n_frames))) - 1) * stride + (k_eff - pad)
    return ideal - length


class CausalConv:
    """CausalConv1d (:135-232).  ``groups`` is either 1 or the channel count (the only two uses in the decoder)."""

    def __init__(self, W, prefix: str, stride: int = 1, dilation: int = 1, depthwise: bool = False):
        self.w = _w(W, prefix + ".conv.weight")
        self.b = _w(W, prefix + ".conv.bias")
        self.stride, self.dilation, self.depthwise = stride, dilation, depthwise
        self.k_eff = (self.w.shape[1] - 1) * dilation + 1
        self.pad = self.k_eff - stride
        self.buffer: Optional[torch.Tensor] = None

    def _apply(self, x: torch.Tensor) -> torch.Tensor:
        if not self.depthwise:
            return conv1d_mlx(x, self.w, self.b, self.stride, self.dilation)
        k = self.w.shape[1]                                        # :186-196: windows * w summed over taps
        n = max(0, x.shape[2] - k + 1)
        y = torch.zeros(x.shape[0], x.shape[1], n, dtype=x.dtype)
        for i in range(k):
            y = y + x[:, :, i: i + n] * self.w[:, i, 0][None, :, None]
        return y + self.b[None, :, None]

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        return self._apply(F.pad(x, (self.pad, extra_padding(x.shape[-1], self.k_eff, self.stride))))

    def step(self, x: torch.Tensor) -> torch.Tensor:              # :199-227
        if self.pad > 0:
            x = torch.cat([self.buffer, x], dim=-1) if self.buffer is not None else F.pad(x, (self.pad, 0))
            self.buffer = x[:, :, max(0, x.shape[2] - self.pad):]
        return self._apply(x)

    def reset(self):
        self.buffer = None


def snake_beta(x: torch.Tensor, alpha: torch.Tensor, beta: torch.Tensor) -> torch.Tensor:
    """SnakeBeta / DecoderOutputSnake (:236-253, :693-708)."""
    a = torch.exp(alpha)[None, :, None]
    b = torch.exp(beta)[None, :, None]
    s = torch.sin(x * a)
    return x + (1.0 / (b + 1e-9)) * s * s


def gelu_exact(x: torch.Tensor) -> torch.Tensor:
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


class ConvNeXt:
    """ConvNeXtBlock (:257-297)."""

    def __init__(self, W, prefix: str):
        self.dw = CausalConv(W, prefix + ".dwconv", depthwise=True)
        self.nw, self.nb = _w(W, prefix + ".norm.weight"), _w(W, prefix + ".norm.bias")
        self.w1, self.b1 = _w(W, prefix + ".pwconv1.weight"), _w(W, prefix + ".pwconv1.bias")
        self.w2, self.b2 = _w(W, prefix + ".pwconv2.weight"), _w(W, prefix + ".pwconv2.bias")
        self.gamma = _w(W, prefix + ".gamma")

    def _tail(self, x, h):
        h = h.transpose(1, 2)
        h = F.layer_norm(h, (h.shape[-1],), self.nw, self.nb, 1e-6)
        h = gelu_exact(h @ self.w1.T + self.b1)
        h = self.gamma * (h @ self.w2.T + self.b2)
        return x + h.transpose(1, 2)

    def __call__(self, x):
        return self._tail(x, self.dw(x))

    def step(self, x):
        return self._tail(x, self.dw.step(x))

    def reset(self):
        self.dw.reset()


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    return w * (x * torch.rsqrt((x * x).mean(-1, keepdim=True) + eps))


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], dim=-1)


def rope_cos_sin(positions: torch.Tensor, head_dim: int, base: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """DecoderRotaryEmbedding (:326-342): cos/sin of [freqs, freqs], shape [T, head_dim]."""
    inv = 1.0 / (base ** (torch.arange(0, head_dim, 2, dtype=DT) / head_dim))
    f = positions.to(DT)[:, None] * inv[None, :]
    e = torch.cat([f, f], dim=-1)
    return torch.cos(e), torch.sin(e)


class PreTransformer:
    """DecoderTransformer (:431-491) with an optional per-layer key/value cache (KVCacheSimple semantics)."""

    def __init__(self, cfg: TokenizerDecoderConfig, W, prefix: str = "pre_transformer"):
        self.cfg, self.W, self.p = cfg, W, prefix

    def make_cache(self) -> List[Optional[Tuple[torch.Tensor, torch.Tensor]]]:
        return [None] * self.cfg.num_hidden_layers

    def _lin(self, name, x):
        y = x @ _w(self.W, name + ".weight").T
        if name + ".bias" in self.W:
            y = y + _w(self.W, name + ".bias")
        return y

    def _layer(self, i, x, cos, sin, mask, cache):
        c, p = self.cfg, f"{self.p}.layers.{i}"
        B, T, _ = x.shape
        h = rms_norm(x, _w(self.W, p + ".input_layernorm.weight"), c.rms_norm_eps)
        q = self._lin(p + ".self_attn.q_proj", h).view(B, T, c.num_attention_heads, c.head_dim).transpose(1, 2)
        k = self._lin(p + ".self_attn.k_proj", h).view(B, T, c.num_key_value_heads, c.head_dim).transpose(1, 2)
        v = self._lin(p + ".self_attn.v_proj", h).view(B, T, c.num_key_value_heads, c.head_dim).transpose(1, 2)
        q = q * cos + rotate_half(q) * sin
        k = k * cos + rotate_half(k) * sin
        if cache is not None:
            if cache[i] is not None:
                k, v = torch.cat([cache[i][0], k], dim=2), torch.cat([cache[i][1], v], dim=2)
            cache[i] = (k, v)
        rep = c.num_attention_heads // c.num_key_value_heads
        kk, vv = k.repeat_interleave(rep, dim=1), v.repeat_interleave(rep, dim=1)
        s = (q @ kk.transpose(-1, -2)) / math.sqrt(c.head_dim)
        if mask is not None:
            s = s + mask
        a = (torch.softmax(s, dim=-1) @ vv).transpose(1, 2).reshape(B, T, -1)
        x = x + _w(self.W, p + ".self_attn_layer_scale.scale") * self._lin(p + ".self_attn.o_proj", a)
        h = rms_norm(x, _w(self.W, p + ".post_attention_layernorm.weight"), c.rms_norm_eps)
        m = self._lin(p + ".mlp.down_proj", F.silu(self._lin(p + ".mlp.gate_proj", h)) * self._lin(p + ".mlp.up_proj", h))
        return x + _w(self.W, p + ".mlp_layer_scale.scale") * m

    def __call__(self, x: torch.Tensor, cache=None) -> torch.Tensor:
        c = self.cfg
        T = x.shape[1]
        x = self._lin(self.p + ".input_proj", x)
        offset = 0 if cache is None or cache[0] is None else cache[0][0].shape[2]
        cos, sin = rope_cos_sin(torch.arange(offset, offset + T), c.head_dim, c.rope_theta)
        mask = None
        if T > 1:                                                  # :463-471: the last T rows of a (offset+T)^2 causal mask
            total = offset + T
            rows = torch.arange(offset, total)[:, None]
            mask = torch.where(torch.arange(total)[None, :] > rows, torch.tensor(-1e9, dtype=DT), torch.tensor(0.0, dtype=DT))
        for i in range(c.num_hidden_layers):
            x = self._layer(i, x, cos, sin, mask, cache)
        return self._lin(self.p + ".output_proj", rms_norm(x, _w(self.W, self.p + ".norm.weight"), c.rms_norm_eps))


class ResidualUnit:
    """DecoderResidualUnit (:495-530)."""

    def __init__(self, W, prefix: str, dilation: int):
        self.a1 = (_w(W, prefix + ".act1.alpha"), _w(W, prefix + ".act1.beta"))
        self.a2 = (_w(W, prefix + ".act2.alpha"), _w(W, prefix + ".act2.beta"))
        self.c1 = CausalConv(W, prefix + ".conv1", dilation=dilation)
        self.c2 = CausalConv(W, prefix + ".conv2")

    def __call__(self, x):
        return x + self.c2(snake_beta(self.c1(snake_beta(x, *self.a1)), *self.a2))

    def step(self, x):
        return x + self.c2.step(snake_beta(self.c1.step(snake_beta(x, *self.a1)), *self.a2))

    def reset(self):
        self.c1.reset(), self.c2.reset()


class BlockUpsample:
    """DecoderBlockUpsample (:533-582)."""

    def __init__(self, W, prefix: str, rate: int):
        self.w, self.b, self.rate = _w(W, prefix + ".conv.weight"), _w(W, prefix + ".conv.bias"), rate
        self.trim = self.w.shape[1] - rate
        self.overflow: Optional[torch.Tensor] = None

    def __call__(self, x):
        h = conv_transpose1d_mlx(x, self.w, self.b, self.rate)
        return h[:, :, : h.shape[2] - self.trim] if self.trim > 0 else h

    def step(self, x):
        h = conv_transpose1d_mlx(x, self.w, self.b, self.rate)
        if self.overflow is not None:
            n = self.overflow.shape[2]
            h = torch.cat([h[:, :, :n] + self.overflow, h[:, :, n:]], dim=-1)
        if self.trim > 0:
            split = max(0, h.shape[2] - self.trim)
            self.overflow, h = h[:, :, split:], h[:, :, :split]
        else:
            self.overflow = None
        return h

    def reset(self):
        self.overflow = None


class DecoderBlock:
    """DecoderBlock (:584-638)."""

    def __init__(self, W, prefix: str, rate: int):
        self.snake = (_w(W, prefix + ".block.0.alpha"), _w(W, prefix + ".block.0.beta"))
        self.up = BlockUpsample(W, prefix + ".block.1", rate)
        self.units = [ResidualUnit(W, f"{prefix}.block.{2 + j}", d) for j, d in enumerate((1, 3, 9))]

    def __call__(self, x):
        x = self.up(snake_beta(x, *self.snake))
        for u in self.units:
            x = u(x)
        return x

    def step(self, x):
        x = self.up.step(snake_beta(x, *self.snake))
        for u in self.units:
            x = u.step(x)
        return x

    def reset(self):
        self.up.reset()
        for u in self.units:
            u.reset()


class EdgeConv:
    """DecoderInitialConv / DecoderOutputConv (:641-683, :711-731): causal k-tap Conv1d with a streaming buffer."""

    def __init__(self, W, prefix: str):
        self.w, self.b = _w(W, prefix + ".conv.weight"), _w(W, prefix + ".conv.bias")
        self.k = self.w.shape[1]
        self.buffer: Optional[torch.Tensor] = None

    def __call__(self, x):
        return conv1d_mlx(F.pad(x, (self.k - 1, 0)), self.w, self.b)

    def step(self, x):
        pad = self.k - 1
        if pad > 0:
            x = torch.cat([self.buffer, x], dim=-1) if self.buffer is not None else F.pad(x, (pad, 0))
            self.buffer = x[:, :, max(0, x.shape[2] - pad):]
        return conv1d_mlx(x, self.w, self.b)

    def reset(self):
        self.buffer = None


class UpsampleLayer:
    """UpsampleLayer (:758-791): CausalTransposeConv1d with k = stride (no trim, stateless) then ConvNeXt."""

    def __init__(self, W, prefix: str, factor: int):
        self.w, self.b, self.factor = _w(W, prefix + ".layers.0.conv.weight"), _w(W, prefix + ".layers.0.conv.bias"), factor
        self.trim = self.w.shape[1] - factor
        self.cn = ConvNeXt(W, prefix + ".layers.1")

    def _up(self, x):
        h = conv_transpose1d_mlx(x, self.w, self.b, self.factor)
        return h[:, :, : h.shape[2] - self.trim] if self.trim > 0 else h

    def __call__(self, x):
        return self.cn(self._up(x))

    def step(self, x):
        return self.cn.step(self._up(x))

    def reset(self):
        self.cn.reset()


# ---------------------------------------------------------------- quantizer

def codebook_embedding(W, prefix: str) -> torch.Tensor:
    """EuclideanCodebook.updateInPlace (Quantization.swift:29-33)."""
    usage = _w(W, prefix + ".cluster_usage").clamp(min=1e-5)
    return _w(W, prefix + ".embedding_sum") / usage[:, None]


def rvq_decode(W, prefix: str, codes: torch.Tensor) -> torch.Tensor:
    """ResidualVectorQuantizer.decode (:79-88): codes [B, n_q, T] -> [B, out_dim, T] (sum of gathers, then the k1 projection)."""
    q = None
    for i in range(codes.shape[1]):
        e = codebook_embedding(W, f"{prefix}.vq.layers.{i}.codebook")[codes[:, i].long()]     # [B, T, D/2]
        q = e if q is None else q + e
    q = q.transpose(1, 2)
    key = prefix + ".output_proj.weight"
    return conv1d_mlx(q, _w(W, key), None) if key in W else q


def quantizer_decode(cfg: TokenizerDecoderConfig, W, codes: torch.Tensor) -> torch.Tensor:
    """SplitResidualVectorQuantizer.decode (:112-119)."""
    s = cfg.num_semantic_quantizers
    q = rvq_decode(W, "quantizer.rvq_first", codes[:, :s])
    if codes.shape[1] > s:
        q = q + rvq_decode(W, "quantizer.rvq_rest", codes[:, s:])
    return q


# ---------------------------------------------------------------- the decoder

class SpeechTokenizerDecoder:
    """Qwen3TTSSpeechTokenizerDecoder (:888-1025).  ``codes`` are ``[B, n_q, T]`` integer arrays; output ``[B, 1, T * total_upsample]``."""

    def __init__(self, cfg: TokenizerDecoderConfig, W: Dict[str, torch.Tensor]):
        self.cfg, self.W = cfg, W
        self.pre_conv = CausalConv(W, "pre_conv")
        self.pre_transformer = PreTransformer(cfg, W)
        self.upsample = [UpsampleLayer(W, f"upsample.{i}", f) for i, f in enumerate(cfg.upsampling_ratios)]
        self.init_conv = EdgeConv(W, "decoder.0")
        self.blocks = [DecoderBlock(W, f"decoder.{1 + b}", r) for b, r in enumerate(cfg.upsample_rates)]
        n = len(cfg.upsample_rates)
        self.out_snake = (_w(W, f"decoder.{n + 1}.alpha"), _w(W, f"decoder.{n + 1}.beta"))
        self.out_conv = EdgeConv(W, f"decoder.{n + 2}")
        self.cache = None

    def __call__(self, codes) -> torch.Tensor:                    # :926-947
        codes = torch.as_tensor(np.asarray(codes))
        h = quantizer_decode(self.cfg, self.W, codes)
        h = self.pre_conv(h)
        h = self.pre_transformer(h.transpose(1, 2)).transpose(1, 2)
        for u in self.upsample:
            h = u(h)
        h = self.init_conv(h)
        for b in self.blocks:
            h = b(h)
        h = self.out_conv(snake_beta(h, *self.out_snake))
        return h.clamp(-1.0, 1.0)

    def reset_streaming_state(self):                              # :949-970
        self.cache = None
        self.pre_conv.reset()
        for u in self.upsample:
            u.reset()
        self.init_conv.reset()
        for b in self.blocks:
            b.reset()
        self.out_conv.reset()

    def streaming_step(self, codes) -> torch.Tensor:              # :973-1008
        codes = torch.as_tensor(np.asarray(codes))
        if self.cache is None:
            self.cache = self.pre_transformer.make_cache()
        h = quantizer_decode(self.cfg, self.W, codes)
        h = self.pre_conv.step(h)
        h = self.pre_transformer(h.transpose(1, 2), cache=self.cache).transpose(1, 2)
        for u in self.upsample:
            h = u.step(h)
        h = self.init_conv.step(h)
        for b in self.blocks:
            h = b.step(h)
        h = self.out_conv.step(snake_beta(h, *self.out_snake))
        return h.clamp(-1.0, 1.0)

    def chunked_decode(self, codes, chunk_size: int = 300, left_context_size: int = 25) -> torch.Tensor:   # :1010-1024
        codes = torch.as_tensor(np.asarray(codes))
        total, start, wavs = codes.shape[-1], 0, []
        while start < total:
            end = min(start + chunk_size, total)
            ctx = left_context_size if start - left_context_size > 0 else start
            w = self(codes[:, :, start - ctx: end])
            wavs.append(w[:, :, ctx * self.cfg.total_upsample:])
            start = end
        return torch.cat(wavs, dim=-1)


def decode(cfg: TokenizerDecoderConfig, W, audio_codes, decode_upsample_rate: Optional[int] = None,
           chunk_size: int = 300, left_context_size: int = 25) -> Tuple[np.ndarray, np.ndarray]:
    """Qwen3TTSSpeechTokenizer.decode (:1059-1068): audio_codes [B, T, n_q] -> (wav [B, samples], valid lengths [B])."""
    ac = np.asarray(audio_codes)
    wav = SpeechTokenizerDecoder(cfg, W).chunked_decode(ac.transpose(0, 2, 1), chunk_size, left_context_size)[:, 0]
    rate = cfg.total_upsample if decode_upsample_rate is None else decode_upsample_rate
    lengths = (ac[:, :, 0] > 0).sum(axis=1).astype(np.int32) * np.int32(rate)
    return wav.numpy(), lengths


def decode_chunk(cfg: TokenizerDecoderConfig, W, audio_codes, chunk_tokens: int = 300, decode_upsample_rate: Optional[int] = None) -> np.ndarray:
    """Qwen3TTSModel.decodeChunk (Qwen3TTS.swift:214-231): the audio the non-streaming generate() returns -- the streaming
    decoder run over ``chunk_tokens``-sized pieces, row 0, cut to (number of frames whose first code > 0) * upsample rate
    (the count is taken over the whole batch, as the reference does)."""
    ac = np.asarray(audio_codes)
    audio = np.concatenate(streaming_decode(cfg, W, ac, chunk_tokens), axis=-1)[0]
    rate = cfg.total_upsample if decode_upsample_rate is None else decode_upsample_rate
    valid = int((ac[:, :, 0] > 0).sum()) * rate
    return audio[:valid] if 0 < valid < audio.shape[0] else audio


def streaming_decode(cfg: TokenizerDecoderConfig, W, audio_codes, chunk_tokens: int = 100) -> List[np.ndarray]:
    """Qwen3TTSSpeechTokenizer.streamingDecode (:1070-1092)."""
    codes = np.asarray(audio_codes).transpose(0, 2, 1)
    d = SpeechTokenizerDecoder(cfg, W)
    d.reset_streaming_state()
    out, start = [], 0
    while start < codes.shape[-1]:
        end = min(start + chunk_tokens, codes.shape[-1])
        out.append(d.streaming_step(codes[:, :, start:end])[:, 0].numpy())
        start = end
    d.reset_streaming_state()
    return out
