"""Oracle for the SNAC codec path.  Test infrastructure only.

Follows:
  Sources/MLXAudioCodecs/SNAC/Layers.swift:36-50    normalizeWeight / snake
  Sources/MLXAudioCodecs/SNAC/Layers.swift:54-118   WNConv1d      (eps 1e-12 in the norm)
  Sources/MLXAudioCodecs/SNAC/Layers.swift:122-183  WNConvTranspose1d (no eps, outputPadding dropped)
  Sources/MLXAudioCodecs/SNAC/Layers.swift:202-232  ResidualUnit
  Sources/MLXAudioCodecs/SNAC/Layers.swift:263-279  NoiseBlock    (noise passed in explicitly here)
  Sources/MLXAudioCodecs/SNAC/Layers.swift:283-315  DecoderBlock
  Sources/MLXAudioCodecs/SNAC/Layers.swift:364-421  Decoder
  Sources/MLXAudioCodecs/SNAC/VQ.swift:14-20,47-120,150-191   normalize / VectorQuantize / RVQ
  Sources/MLXAudioCodecs/SNAC/SNACDecoder.swift:127-131        SNAC.decode

Weights live in a flat ``dict`` keyed exactly like the reference's safetensors
(``decoder.model.layers.2.block.layers.1.weight_v`` ...), MLX layouts:
Conv1d ``[out, k, in/groups]``, ConvTranspose1d ``weight_v [in, k, out]``.
MLX ``conv1d`` is NLC cross-correlation; ``convTransposed1d`` is taken to have
scatter semantics ``y[t*s + k - pad] += x[t] * w[o,k,i]`` (SURVEY.md 8c trap 7).
The signal path is float64 (the "ideal" value); ``quantize_fp32`` additionally
restates the RVQ encode arithmetic in explicitly ordered float32 so that code
indices can be compared bit-exactly.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class SNACConfig:
    """Sources/MLXAudioCodecs/SNAC/Config.swift:11-37 (defaults = snac_24khz)."""
    sampling_rate: int = 24000
    encoder_dim: int = 48
    encoder_rates: Sequence[int] = (2, 4, 8, 8)
    latent_dim: Optional[int] = None
    decoder_dim: int = 1024
    decoder_rates: Sequence[int] = (8, 8, 4, 2)
    attn_window_size: Optional[int] = None
    codebook_size: int = 4096
    codebook_dim: int = 8
    vq_strides: Sequence[int] = (4, 2, 1)
    noise: bool = True
    depthwise: bool = True

    @property
    def latent(self) -> int:                      # SNACDecoder.swift:50-51
        return self.latent_dim or self.encoder_dim * 2 ** len(self.encoder_rates)

    @property
    def hop_length(self) -> int:                  # SNACDecoder.swift:54
        return int(np.prod(self.encoder_rates))


def init_weights(cfg: SNACConfig, seed: int = 1234) -> Dict[str, np.ndarray]:
    """Random-init decoder + quantizer weights: conv v ~ U(+-1/sqrt(fan_in*k))
    as Layers.swift:81-86, g = ||v|| perturbed, small random biases, Snake alpha ~ U(0.5,1.5)."""
    rng = np.random.default_rng(seed)
    w: Dict[str, np.ndarray] = {}

    def wn(prefix, shape, fan, bias_n, except_last=False):
        s = math.sqrt(1.0 / fan)
        v = rng.uniform(-s, s, size=shape).astype(np.float32)
        g = np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True))
        g = (g * rng.uniform(0.8, 1.2, size=g.shape)).astype(np.float32)
        w[prefix + ".weight_v"], w[prefix + ".weight_g"] = v, g
        if bias_n:
            w[prefix + ".bias"] = rng.uniform(-0.05, 0.05, size=bias_n).astype(np.float32)

    D, C = cfg.latent, cfg.decoder_dim
    for i, _ in enumerate(cfg.vq_strides):
        q = f"quantizer.quantizers.{i}"
        wn(q + ".in_proj", (cfg.codebook_dim, 1, D), D, cfg.codebook_dim)
        wn(q + ".out_proj", (D, 1, cfg.codebook_dim), cfg.codebook_dim, D)
        w[q + ".codebook.weight"] = rng.standard_normal((cfg.codebook_size, cfg.codebook_dim)).astype(np.float32)
    p = "decoder.model.layers"
    li = 0
    if cfg.depthwise:
        wn(f"{p}.0", (D, 7, 1), 7 * D, D)          # fan uses inChannels*k as Layers.swift:81
        wn(f"{p}.1", (C, 1, D), D, C)
        li = 2
    else:
        wn(f"{p}.0", (C, 7, D), 7 * D, C)
        li = 1
    for i, s in enumerate(cfg.decoder_rates):
        cin, cout = C // 2 ** i, C // 2 ** (i + 1)
        b = f"{p}.{li}.block.layers"
        w[f"{b}.0.alpha"] = rng.uniform(0.5, 1.5, size=(1, cin, 1)).astype(np.float32)
        wn(f"{b}.1", (cin, 2 * s, cout), cin * 2 * s, cout)
        j = 2
        if cfg.noise:
            wn(f"{b}.2.linear", (cout, 1, cout), cout, 0)
            j = 3
        for _dil in (1, 3, 9):
            r = f"{b}.{j}.block.layers"
            g = cout if cfg.depthwise else 1
            w[f"{r}.0.alpha"] = rng.uniform(0.5, 1.5, size=(1, cout, 1)).astype(np.float32)
            wn(f"{r}.1", (cout, 7, cout // g), cout * 7, cout)
            w[f"{r}.2.alpha"] = rng.uniform(0.5, 1.5, size=(1, cout, 1)).astype(np.float32)
            wn(f"{r}.3", (cout, 1, cout), cout, cout)
            j += 1
        li += 1
    cf = C // 2 ** len(cfg.decoder_rates)
    w[f"{p}.{li}.alpha"] = rng.uniform(0.5, 1.5, size=(1, cf, 1)).astype(np.float32)
    wn(f"{p}.{li + 1}", (1, 7, cf), cf * 7, 1)
    return w


# --------------------------------------------------------------------------- primitives

DTYPE = torch.float64   # bench.py's cpu_baseline leg switches this to float32 (the reference computes in fp32)


def _t(a) -> torch.Tensor:
    return torch.as_tensor(np.asarray(a), dtype=DTYPE)


def wn_conv_weight(w: Dict, prefix: str) -> torch.Tensor:
    """Layers.swift:102-103: g * v / (||v||_{axes 1,2} + 1e-12); MLX [out,k,in/g] -> torch [out,in/g,k]."""
    v, g = _t(w[prefix + ".weight_v"]), _t(w[prefix + ".weight_g"])
    nrm = torch.sqrt((v ** 2).sum(dim=(1, 2), keepdim=True))
    return (g * v / (nrm + 1e-12)).permute(0, 2, 1).contiguous()


def wn_convT_weight(w: Dict, prefix: str) -> torch.Tensor:
    """Layers.swift:166-168: g * v / ||v||_{axes 1,2} (no eps) on [in,k,out]; -> torch [in,out,k]."""
    v, g = _t(w[prefix + ".weight_v"]), _t(w[prefix + ".weight_g"])
    nrm = torch.sqrt((v ** 2).sum(dim=(1, 2), keepdim=True))
    return (g * v / nrm).permute(0, 2, 1).contiguous()


def snake(x: torch.Tensor, alpha: torch.Tensor) -> torch.Tensor:
    """Layers.swift:44-50: x + 1/(alpha+1e-9) * sin(alpha*x)^2, alpha [1,C,1]."""
    return x + (1.0 / (alpha + 1e-9)) * torch.sin(alpha * x) ** 2


def wn_conv1d(w, prefix, x, *, padding=0, dilation=1, groups=1, stride=1):
    b = w.get(prefix + ".bias")
    return F.conv1d(x, wn_conv_weight(w, prefix), None if b is None else _t(b), stride=stride,
                    padding=padding, dilation=dilation, groups=groups)


def wn_conv_transpose1d(w, prefix, x, *, stride, padding):
    b = w.get(prefix + ".bias")
    return F.conv_transpose1d(x, wn_convT_weight(w, prefix), None if b is None else _t(b),
                              stride=stride, padding=padding)


def residual_unit(w, prefix, x, dilation, groups):
    """Layers.swift:202-232."""
    p = prefix + ".block.layers"
    y = snake(x, _t(w[p + ".0.alpha"]))
    y = wn_conv1d(w, p + ".1", y, padding=(6 * dilation) // 2, dilation=dilation, groups=groups)
    y = snake(y, _t(w[p + ".2.alpha"]))
    y = wn_conv1d(w, p + ".3", y)
    return x + y


def from_codes(cfg: SNACConfig, w: Dict, codes: List[np.ndarray]) -> torch.Tensor:
    """VQ.swift:165-191: sum_i repeat_interleave(outProj_i(codebook_i[codes_i]^T), stride_i) -> [B,D,T]."""
    z = 0.0
    for i, s in enumerate(cfg.vq_strides):
        q = f"quantizer.quantizers.{i}"
        emb = _t(w[q + ".codebook.weight"])[torch.as_tensor(np.asarray(codes[i]), dtype=torch.long)]
        zi = wn_conv1d(w, q + ".out_proj", emb.transpose(1, 2))
        if s > 1:
            zi = torch.repeat_interleave(zi, s, dim=2)
        z = z + zi
    return z


def decoder(cfg: SNACConfig, w: Dict, z: torch.Tensor, noise: Optional[List[np.ndarray]]) -> torch.Tensor:
    """Layers.swift:364-421.  ``noise[i]`` is the [B,1,T_i] Gaussian tensor NoiseBlock i would draw
    (Layers.swift:274); ``None`` entries / ``noise=None`` mean a zero draw."""
    assert cfg.attn_window_size is None, "LocalMHA only exists in the 32/44 kHz models (SURVEY 8c trap 10)"
    p = "decoder.model.layers"
    D = cfg.latent
    if cfg.depthwise:
        x = wn_conv1d(w, f"{p}.0", z, padding=3, groups=D)
        x = wn_conv1d(w, f"{p}.1", x)
        li = 2
    else:
        x = wn_conv1d(w, f"{p}.0", z, padding=3)
        li = 1
    for i, s in enumerate(cfg.decoder_rates):
        cout = cfg.decoder_dim // 2 ** (i + 1)
        b = f"{p}.{li}.block.layers"
        x = snake(x, _t(w[f"{b}.0.alpha"]))
        x = wn_conv_transpose1d(w, f"{b}.1", x, stride=s, padding=math.ceil(s / 2))
        j = 2
        if cfg.noise:
            h = wn_conv1d(w, f"{b}.2.linear", x)
            if noise is not None and noise[i] is not None:
                x = x + _t(noise[i]) * h
            j = 3
        for dil in (1, 3, 9):
            x = residual_unit(w, f"{b}.{j}", x, dil, cout if cfg.depthwise else 1)
            j += 1
        li += 1
    x = snake(x, _t(w[f"{p}.{li}.alpha"]))
    x = wn_conv1d(w, f"{p}.{li + 1}", x, padding=3)
    return torch.tanh(x)


def decode(cfg: SNACConfig, w: Dict, codes: List[np.ndarray],
           noise: Optional[List[np.ndarray]] = None) -> np.ndarray:
    """SNACDecoder.swift:127-131 -> waveform [B,1,T*hop] float64."""
    with torch.no_grad():
        return decoder(cfg, w, from_codes(cfg, w, codes), noise).numpy()


def noise_shapes(cfg: SNACConfig, batch: int, t_latent: int) -> List[tuple]:
    out, t = [], t_latent
    for s in cfg.decoder_rates:
        t *= s
        out.append((batch, 1, t))
    return out


# --------------------------------------------------------------------------- RVQ encode side (a16)

def _f32(x):
    return np.asarray(x, dtype=np.float32)


def l2_normalize_fp32(x: np.ndarray) -> np.ndarray:
    """VQ.swift:14-20 in ordered float32: x / max(sqrt(sum_d x_d^2 sequential), 1e-12)."""
    x = _f32(x)
    acc = np.zeros(x.shape[0], dtype=np.float32)
    for d in range(x.shape[1]):
        acc = (acc + x[:, d] * x[:, d]).astype(np.float32)
    nrm = np.maximum(np.sqrt(acc, dtype=np.float32), np.float32(1e-12))
    return (x / nrm[:, None]).astype(np.float32)


def nearest_code_fp32(enc: np.ndarray, codebook: np.ndarray) -> np.ndarray:
    """VQ.swift:96-120 decodeLatents in explicitly ordered float32 (no FMA contraction):
    dist[n] = (|e|^2 - 2*(e.c_n)) + |c_n|^2 with sequential d-loops; index = first max of -dist."""
    e, c = l2_normalize_fp32(enc), l2_normalize_fp32(codebook)
    D = e.shape[1]
    e2 = np.zeros(e.shape[0], np.float32)
    c2 = np.zeros(c.shape[0], np.float32)
    for d in range(D):
        e2 = (e2 + e[:, d] * e[:, d]).astype(np.float32)
        c2 = (c2 + c[:, d] * c[:, d]).astype(np.float32)
    out = np.empty(e.shape[0], np.int32)
    for r0 in range(0, e.shape[0], 512):
        eb = e[r0:r0 + 512]
        dot = np.zeros((eb.shape[0], c.shape[0]), np.float32)
        for d in range(D):
            dot = (dot + (eb[:, d:d + 1] * c[None, :, d]).astype(np.float32)).astype(np.float32)
        dist = ((e2[r0:r0 + 512, None] - (np.float32(2) * dot).astype(np.float32)).astype(np.float32)
                + c2[None, :]).astype(np.float32)
        out[r0:r0 + 512] = np.argmax(-dist, axis=1)
    return out


def quantize(cfg: SNACConfig, w: Dict, z: np.ndarray, fp32_search: bool = True):
    """VQ.swift:150-163 + 47-86: residual VQ of latent z [B,D,T] -> (z_q, [codes_i [B,T/stride_i]]).

    avg-pool(stride) -> in_proj -> nearest code -> gather -> out_proj -> repeat -> residual update.
    With ``fp32_search`` the projection result is rounded to float32 and searched by
    ``nearest_code_fp32`` (the bit-exact definition used for index parity)."""
    with torch.no_grad():
        zt = _t(z)
        residual = zt.clone()
        zq = torch.zeros_like(zt)
        codes = []
        for i, s in enumerate(cfg.vq_strides):
            q = f"quantizer.quantizers.{i}"
            x = residual
            if s > 1:
                x = F.avg_pool1d(x, kernel_size=s, stride=s)
            ze = wn_conv1d(w, q + ".in_proj", x)                              # [B,8,T/s]
            B, Dc, Ts = ze.shape
            enc = ze.permute(0, 2, 1).reshape(B * Ts, Dc).numpy()
            cb = np.asarray(w[q + ".codebook.weight"])
            if fp32_search:
                idx = nearest_code_fp32(enc.astype(np.float32), cb)
            else:
                e = enc / np.maximum(np.linalg.norm(enc, axis=1, keepdims=True), 1e-12)
                c = cb.astype(np.float64)
                c = c / np.maximum(np.linalg.norm(c, axis=1, keepdims=True), 1e-12)
                dist = (e ** 2).sum(1, keepdims=True) - 2 * e @ c.T + (c ** 2).sum(1)[None]
                idx = np.argmax(-dist, axis=1).astype(np.int32)
            idx = idx.reshape(B, Ts)
            zqi = _t(cb)[torch.as_tensor(idx, dtype=torch.long)].transpose(1, 2)
            zqi = wn_conv1d(w, q + ".out_proj", zqi)
            if s > 1:
                zqi = torch.repeat_interleave(zqi, s, dim=2)
            zq = zq + zqi
            residual = residual - zqi
            codes.append(idx.astype(np.int32))
        return zq.numpy(), codes


def synth_codes(cfg: SNACConfig, batch: int, t_finest: int, seed: int = 2) -> List[np.ndarray]:
    """uniform codes, seed 2; level i has t_finest*min(strides)/stride_i entries."""
    rng = np.random.default_rng(seed)
    return [rng.integers(0, cfg.codebook_size, size=(batch, t_finest // s), dtype=np.int32)
            for s in cfg.vq_strides]
