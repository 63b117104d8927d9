"""Oracle for the Vocos vocoder decode path.  Test infrastructure only.

Follows:
  Sources/MLXAudioCodecs/Vocos/VocosBackbone.swift:18-100    ConvNeXtBlock (dw k7 -> LayerNorm(1e-6) -> Linear -> GELU -> Linear -> gamma, +x)
  Sources/MLXAudioCodecs/Vocos/VocosBackbone.swift:109-204   VocosBackbone (embed conv -> LayerNorm -> blocks -> final LayerNorm)
  Sources/MLXAudioCodecs/Vocos/Vocos.swift:54-179            ISTFTHead (Linear -> exp/clip(1e2) magnitude, phase -> irfft ->
                                                             SYMMETRIC Hann -> overlap-add / window-SUM, trim n_fft/2)
  Sources/MLXAudioCodecs/Vocos/Vocos.swift:284-322           Vocos.decode / decodeAudio
  Sources/MLXAudioCodecs/Vocos/Vocos.swift:17-47             AdaLayerNorm (adanorm_num_embeddings > 0): parameter-free LayerNorm(1e-6),
                                                             then * scale(cond)[:, None, :] + shift(cond)[:, None, :] with scale / shift
                                                             Linear(num_embeddings -> dim) of the conditioning row (`bandwidthId`)
Weights use the reference's safetensors keys and MLX layouts (Conv1d ``[out, k, in/groups]``, Linear ``[out, in]``).
float64 signal path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class VocosConfig:
    """Test geometry of the reference (Tests/MLXAudioCodecsTests.swift:419-431): dim 512, 8 layers, n_fft 1024, hop 256."""
    input_channels: int = 100
    dim: int = 512
    intermediate_dim: int = 1536
    num_layers: int = 8
    n_fft: int = 1024
    hop_length: int = 256
    input_kernel_size: int = 7
    dw_kernel_size: int = 7
    adanorm_num_embeddings: int = 0        # > 0: every backbone norm except the final one is an AdaLayerNorm


def init_weights(cfg: VocosConfig, seed: int = 1234) -> Dict[str, np.ndarray]:
    rng = np.random.default_rng(seed)
    w: Dict[str, np.ndarray] = {}

    def u(shape, fan):
        s = (1.0 / fan) ** 0.5
        return rng.uniform(-s, s, size=shape).astype(np.float32)

    d, I = cfg.dim, cfg.intermediate_dim
    w["backbone.embed.weight"] = u((d, cfg.input_kernel_size, cfg.input_channels), cfg.input_channels * cfg.input_kernel_size)
    w["backbone.embed.bias"] = u((d,), d)
    E = cfg.adanorm_num_embeddings

    def norm(p, ada):
        if ada and E > 0:
            w[p + ".scale.weight"] = (1.0 / E + 0.3 * rng.standard_normal((d, E))).astype(np.float32)
            w[p + ".scale.bias"] = (0.1 * rng.standard_normal(d)).astype(np.float32)
            w[p + ".shift.weight"] = (0.2 * rng.standard_normal((d, E))).astype(np.float32)
            w[p + ".shift.bias"] = (0.05 * rng.standard_normal(d)).astype(np.float32)
        else:
            w[p + ".weight"] = (1.0 + 0.1 * rng.standard_normal(d)).astype(np.float32)
            w[p + ".bias"] = (0.05 * rng.standard_normal(d)).astype(np.float32)

    norm("backbone.norm", True)
    norm("backbone.final_layer_norm", False)
    for l in range(cfg.num_layers):
        p = f"backbone.convnext.{l}."
        w[p + "dwconv.weight"] = u((d, cfg.dw_kernel_size, 1), cfg.dw_kernel_size)
        w[p + "dwconv.bias"] = u((d,), d)
        norm(p + "norm", True)
        w[p + "pwconv1.weight"] = u((I, d), d)
        w[p + "pwconv1.bias"] = u((I,), d)
        w[p + "pwconv2.weight"] = u((d, I), I)
        w[p + "pwconv2.bias"] = u((d,), I)
        w[p + "gamma"] = (1.0 / cfg.num_layers * (1.0 + 0.2 * rng.standard_normal(d))).astype(np.float32)
    w["head.out.weight"] = (0.3 * u((cfg.n_fft + 2, d), d)).astype(np.float32)
    w["head.out.bias"] = (0.05 * rng.standard_normal(cfg.n_fft + 2)).astype(np.float32)
    return w


def _t(a):
    return torch.as_tensor(np.asarray(a), dtype=torch.float64)


def hann_symmetric(n: int) -> torch.Tensor:
    """Vocos.swift:170-178: 0.5 - 0.5*cos(2*pi*i/(n-1))."""
    i = torch.arange(n, dtype=torch.float64)
    return 0.5 - 0.5 * torch.cos(2.0 * np.pi * i / (n - 1))


def _norm(cfg: VocosConfig, w: Dict, p: str, y: torch.Tensor, cond) -> torch.Tensor:
    """LayerNorm(1e-6), or AdaLayerNorm (Vocos.swift:31-46) when the model has one at `p`: cond [B, num_embeddings]."""
    d = cfg.dim
    if cfg.adanorm_num_embeddings > 0 and p + ".scale.weight" in w:
        if cond is None:
            raise ValueError("AdaLayerNorm requires bandwidthId")          # the reference fatalErrors (VocosBackbone.swift:66-68,181-183)
        c = _t(cond)
        scale = c @ _t(w[p + ".scale.weight"]).T + _t(w[p + ".scale.bias"])
        shift = c @ _t(w[p + ".shift.weight"]).T + _t(w[p + ".shift.bias"])
        mean = y.mean(-1, keepdim=True)
        var = y.var(-1, unbiased=False, keepdim=True)
        return (y - mean) / torch.sqrt(var + 1e-6) * scale[:, None, :] + shift[:, None, :]
    return F.layer_norm(y, (d,), _t(w[p + ".weight"]), _t(w[p + ".bias"]), 1e-6)


@torch.no_grad()
def convnext_layer(cfg: VocosConfig, w: Dict, l: int, h: torch.Tensor, cond=None) -> torch.Tensor:
    """ConvNeXtBlock l on h [B, L, dim] (VocosBackbone.swift:18-100): depthwise k "same" -> LayerNorm -> Linear -> exact GELU -> Linear -> gamma, + h."""
    d = cfg.dim
    p = f"backbone.convnext.{l}."
    y = F.conv1d(h.transpose(1, 2), _t(w[p + "dwconv.weight"]).permute(0, 2, 1), _t(w[p + "dwconv.bias"]),
                 padding=cfg.dw_kernel_size // 2, groups=d).transpose(1, 2)
    y = _norm(cfg, w, p + "norm", y, cond)
    y = F.gelu(y @ _t(w[p + "pwconv1.weight"]).T + _t(w[p + "pwconv1.bias"]))
    y = y @ _t(w[p + "pwconv2.weight"]).T + _t(w[p + "pwconv2.bias"])
    return h + _t(w[p + "gamma"]) * y


@torch.no_grad()
def backbone(cfg: VocosConfig, w: Dict, feats: np.ndarray, cond=None) -> torch.Tensor:
    """feats [B, L, input_channels] -> [B, L, dim]."""
    x = _t(feats).transpose(1, 2)
    h = F.conv1d(x, _t(w["backbone.embed.weight"]).permute(0, 2, 1), _t(w["backbone.embed.bias"]), padding=cfg.input_kernel_size // 2)
    h = h.transpose(1, 2)
    d = cfg.dim
    h = _norm(cfg, w, "backbone.norm", h, cond)
    for l in range(cfg.num_layers):
        h = convnext_layer(cfg, w, l, h, cond)
    return F.layer_norm(h, (d,), _t(w["backbone.final_layer_norm.weight"]), _t(w["backbone.final_layer_norm.bias"]), 1e-6)


@torch.no_grad()
def istft_head(cfg: VocosConfig, w: Dict, x: torch.Tensor) -> np.ndarray:
    """x [B, L, dim] -> audio [B, (L-1)*hop] (Vocos.swift:68-167)."""
    h = x @ _t(w["head.out.weight"]).T + _t(w["head.out.bias"])       # [B, L, n_fft+2]
    half = (cfg.n_fft + 2) // 2
    mag = torch.clamp(torch.exp(h[..., :half]), max=1e2)
    ph = h[..., half:]
    spec = torch.complex(mag * torch.cos(ph), mag * torch.sin(ph))    # [B, L, n_fft/2+1]
    frames = torch.fft.irfft(spec, n=cfg.n_fft, dim=-1)               # [B, L, n_fft]
    win = hann_symmetric(cfg.n_fft)
    frames = frames * win
    B, L, _ = frames.shape
    out_len = (L - 1) * cfg.hop_length + cfg.n_fft
    audio = torch.zeros(B, out_len, dtype=torch.float64)
    wsum = torch.zeros(out_len, dtype=torch.float64)
    for i in range(L):
        s = i * cfg.hop_length
        audio[:, s:s + cfg.n_fft] += frames[:, i]
        wsum[s:s + cfg.n_fft] += win
    nz = wsum != 0
    audio[:, nz] = audio[:, nz] / wsum[nz]
    a, b = cfg.n_fft // 2, out_len - cfg.n_fft // 2
    return (audio[:, a:b] if b > a else audio).numpy()


def decode(cfg: VocosConfig, w: Dict, feats: np.ndarray, cond=None) -> np.ndarray:
    """Vocos.decode (Vocos.swift:302-306); cond = the `bandwidthId` rows [B, num_embeddings] of an AdaLayerNorm model."""
    return istft_head(cfg, w, backbone(cfg, w, feats, cond))
