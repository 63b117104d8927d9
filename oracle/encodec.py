"""Oracle for the Encodec decode path.  Test infrastructure only.

Follows (paths relative to the reference checkout):
  Sources/MLXAudioCodecs/Encodec/EncodecQuantization.swift:117-133   EncodecResidualVectorQuantizer.decode (sum of gathers)
  Sources/MLXAudioCodecs/Encodec/EncodecLayers.swift:15-66           EncodecLSTM (gate order i, f, g, o; bias on the x side)
  Sources/MLXAudioCodecs/Encodec/EncodecLayers.swift:72-88           EncodecLSTMBlock (stack + skip)
  Sources/MLXAudioCodecs/Encodec/EncodecLayers.swift:92-211          EncodecConv1d (causal / asymmetric padding, reflect or zero)
  Sources/MLXAudioCodecs/Encodec/EncodecLayers.swift:216-273         EncodecConvTranspose1dLayer (trim paddingTotal)
  Sources/MLXAudioCodecs/Encodec/EncodecLayers.swift:278-337         EncodecResnetBlock (ELU-conv-ELU-conv + conv shortcut)
  Sources/MLXAudioCodecs/Encodec/EncodecLayers.swift:371-450         EncodecBaseConvTranspose1d (scatter y[t*s+k] += x[t] w[oc,k,ic])
  Sources/MLXAudioCodecs/Encodec/Encodec.swift:94-167                EncodecDecoder (layer order)
  Sources/MLXAudioCodecs/Encodec/Encodec.swift:294-402               decodeFrame / linearOverlapAdd / decode
Only norm_type "weight_norm" is restated: in the reference that variant holds plain (already folded) conv weights and no
norm layer (EncodecLayers.swift:133-137); "time_group_norm" (the 48 kHz stereo model) is not on the BASELINE path.
Weights use the reference's parameter keys and MLX layouts (Conv1d / ConvTranspose1d ``[out, k, in]``, LSTM ``Wx [4H, in]``,
``Wh [4H, H]``, ``bias [4H]``, codebooks ``[size, dim]``).  float64 signal path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np


@dataclass
class EncodecConfig:
    """Defaults = EncodecConfig.swift:116-141 (the 24 kHz mono model)."""
    audio_channels: int = 1
    num_filters: int = 32
    kernel_size: int = 7
    num_residual_layers: int = 1
    dilation_growth_rate: int = 2
    codebook_size: int = 1024
    codebook_dim: int = 128
    hidden_size: int = 128
    num_lstm_layers: int = 2
    residual_kernel_size: int = 3
    use_causal_conv: bool = True
    normalize: bool = False
    pad_mode: str = "reflect"
    norm_type: str = "weight_norm"
    last_kernel_size: int = 7
    trim_right_ratio: float = 1.0
    compress: int = 2
    upsampling_ratios: List[int] = field(default_factory=lambda: [8, 5, 4, 2])
    target_bandwidths: List[float] = field(default_factory=lambda: [1.5, 3.0, 6.0, 12.0, 24.0])
    sampling_rate: int = 24000
    chunk_length_s: Optional[float] = None
    overlap: Optional[float] = None
    use_conv_shortcut: bool = True

    @property
    def hop_length(self) -> int:
        return int(np.prod(self.upsampling_ratios))

    @property
    def frame_rate(self) -> int:                 # EncodecQuantization.swift:81-82
        return math.ceil(self.sampling_rate / self.hop_length)

    @property
    def num_quantizers(self) -> int:             # EncodecQuantization.swift:84-85
        return int(1000 * max(self.target_bandwidths) / (self.frame_rate * 10))

    @property
    def chunk_length(self) -> Optional[int]:     # Encodec.swift:196-201
        return None if self.chunk_length_s is None else int(self.chunk_length_s * self.sampling_rate)

    @property
    def chunk_stride(self) -> Optional[int]:     # Encodec.swift:203-208
        if self.chunk_length_s is None or self.overlap is None:
            return None
        return max(1, int((1.0 - self.overlap) * self.chunk_length))


def decoder_layout(cfg: EncodecConfig):
    """The decoder's `layers` list (Encodec.swift:97-148) as (index, kind, params).  Indices are the positions in the
    reference's module array (ELU modules occupy slots too), which is what the weight keys use."""
    out = []
    scaling = 2 ** len(cfg.upsampling_ratios)
    i = 0
    out.append((i, "conv", dict(cin=cfg.hidden_size, cout=scaling * cfg.num_filters, k=cfg.kernel_size, dilation=1))); i += 1
    out.append((i, "lstm", dict(dim=scaling * cfg.num_filters))); i += 1
    for ratio in cfg.upsampling_ratios:
        cur = scaling * cfg.num_filters
        out.append((i, "elu", {})); i += 1
        out.append((i, "convt", dict(cin=cur, cout=cur // 2, k=2 * ratio, stride=ratio))); i += 1
        for j in range(cfg.num_residual_layers):
            out.append((i, "resnet", dict(dim=cur // 2, dilations=[cfg.dilation_growth_rate ** j, 1]))); i += 1
        scaling //= 2
    out.append((i, "elu", {})); i += 1
    out.append((i, "conv", dict(cin=cfg.num_filters, cout=cfg.audio_channels, k=cfg.last_kernel_size, dilation=1))); i += 1
    return out


def init_weights(cfg: EncodecConfig, seed: int = 1234, n_codebooks: Optional[int] = None) -> Dict[str, np.ndarray]:
    """Random-init decoder + codebooks, U(+-1/sqrt(fan_in)) convs / LSTM (PyTorch's default, what the checkpoints were
    trained from), N(0,1) codebooks."""
    rng = np.random.default_rng(seed)
    w: Dict[str, np.ndarray] = {}

    def u(shape, fan):
        s = (1.0 / fan) ** 0.5
        return rng.uniform(-s, s, size=shape).astype(np.float32)

    nq = cfg.num_quantizers if n_codebooks is None else n_codebooks
    for q in range(nq):
        w[f"quantizer.layers.{q}.codebook.embed"] = rng.standard_normal((cfg.codebook_size, cfg.codebook_dim)).astype(np.float32)
    for idx, kind, p in decoder_layout(cfg):
        pre = f"decoder.layers.{idx}."
        if kind == "conv":
            w[pre + "conv.weight"] = u((p["cout"], p["k"], p["cin"]), p["k"] * p["cin"])
            w[pre + "conv.bias"] = u((p["cout"],), p["k"] * p["cin"])
        elif kind == "convt":
            w[pre + "conv.weight"] = u((p["cout"], p["k"], p["cin"]), p["k"] * p["cin"])
            w[pre + "conv.bias"] = u((p["cout"],), p["k"] * p["cin"])
        elif kind == "lstm":
            d = p["dim"]
            for l in range(cfg.num_lstm_layers):
                w[pre + f"lstm.{l}.Wx"] = u((4 * d, d), d)
                w[pre + f"lstm.{l}.Wh"] = u((4 * d, d), d)
                w[pre + f"lstm.{l}.bias"] = u((4 * d,), d)
        elif kind == "resnet":
            dim, hid = p["dim"], p["dim"] // cfg.compress
            ks = [cfg.residual_kernel_size, 1]
            for bi, (k, _d) in enumerate(zip(ks, p["dilations"])):
                cin = dim if bi == 0 else hid
                cout = dim if bi == len(ks) - 1 else hid
                # block = [ELU, conv, ELU, conv] -> conv slots 1 and 3 (EncodecLayers.swift:293-305)
                w[pre + f"block.{2 * bi + 1}.conv.weight"] = u((cout, k, cin), k * cin)
                w[pre + f"block.{2 * bi + 1}.conv.bias"] = u((cout,), k * cin)
            if cfg.use_conv_shortcut:
                w[pre + "shortcut.conv.weight"] = u((dim, 1, dim), dim)
                w[pre + "shortcut.conv.bias"] = u((dim,), dim)
    return w


# ------------------------------------------------------------------ layers (float64, channels-last [B, T, C])

def elu(x: np.ndarray) -> np.ndarray:
    """EncodecLayers.swift:352-365, alpha = 1."""
    return np.where(x > 0, x, np.expm1(np.minimum(x, 0.0)))


def pad1d(x: np.ndarray, left: int, right: int, mode: str) -> np.ndarray:
    """EncodecLayers.swift:147-189.  Reflect indices are clamped, so a pad longer than the signal repeats the edge sample."""
    if mode != "reflect":
        return np.pad(x, ((0, 0), (left, right), (0, 0)))
    L = x.shape[1]
    li = [min(left - i, L - 1) for i in range(left)]
    ri = [max(L - 2 - i, 0) for i in range(right)]
    return np.concatenate([x[:, li, :], x, x[:, ri, :]], axis=1)


def conv1d(cfg: EncodecConfig, x: np.ndarray, w: np.ndarray, b: np.ndarray, stride: int = 1, dilation: int = 1) -> np.ndarray:
    """EncodecConv1d.callAsFunction (EncodecLayers.swift:191-211).  w [out, k, in].  Note padding_total = k - stride ignores
    the dilation (:117) while the extra padding uses the effective kernel size (:116,139-145)."""
    cout, k, cin = w.shape
    L = x.shape[1]
    k_eff = (k - 1) * dilation + 1
    padding_total = k - stride
    n_frames = (L - k_eff + padding_total) / stride + 1
    ideal = (math.ceil(n_frames) - 1) * stride + k_eff - padding_total
    extra = max(0, ideal - L)
    if cfg.use_causal_conv:
        xp = pad1d(x, padding_total, extra, cfg.pad_mode)
    else:
        pr = padding_total // 2
        xp = pad1d(x, padding_total - pr, pr + extra, cfg.pad_mode)
    Lp = xp.shape[1]
    Lo = (Lp - k_eff) // stride + 1
    y = np.zeros((x.shape[0], Lo, cout), dtype=np.float64)
    w64 = w.astype(np.float64)
    for kk in range(k):
        seg = xp[:, kk * dilation: kk * dilation + (Lo - 1) * stride + 1: stride, :]
        y += seg @ w64[:, kk, :].T
    return y + b.astype(np.float64)


def conv_transpose1d(cfg: EncodecConfig, x: np.ndarray, w: np.ndarray, b: np.ndarray, stride: int) -> np.ndarray:
    """EncodecBaseConvTranspose1d (:371-450, y[t*s + k] += x[t] . w[oc, k, :]) + the trim of EncodecConvTranspose1dLayer (:253-272)."""
    cout, k, cin = w.shape
    B, L, _ = x.shape
    Lo = (L - 1) * stride + k
    y = np.zeros((B, Lo, cout), dtype=np.float64)
    w64 = w.astype(np.float64)
    for kk in range(k):
        y[:, kk: kk + (L - 1) * stride + 1: stride, :] += x @ w64[:, kk, :].T
    y += b.astype(np.float64)
    padding_total = k - stride
    pr = math.ceil(padding_total * cfg.trim_right_ratio) if cfg.use_causal_conv else padding_total // 2
    pl = padding_total - pr
    end = Lo - pr
    return y[:, pl:end, :] if end > pl else y


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def lstm(x: np.ndarray, Wx: np.ndarray, Wh: np.ndarray, bias: np.ndarray) -> np.ndarray:
    """EncodecLSTM.callAsFunction (:27-65): zero initial state, gates split i | f | g | o."""
    B, T, _ = x.shape
    H = Wh.shape[1]
    xp = x @ Wx.astype(np.float64).T + bias.astype(np.float64)
    Wh64 = Wh.astype(np.float64).T
    h = np.zeros((B, H)); c = np.zeros((B, H))
    out = np.empty((B, T, H), dtype=np.float64)
    for t in range(T):
        g = xp[:, t, :] + h @ Wh64
        i, f, gg, o = _sigmoid(g[:, :H]), _sigmoid(g[:, H:2 * H]), np.tanh(g[:, 2 * H:3 * H]), _sigmoid(g[:, 3 * H:])
        c = f * c + i * gg
        h = o * np.tanh(c)
        out[:, t, :] = h
    return out


def lstm_block(cfg: EncodecConfig, W: Dict[str, np.ndarray], pre: str, x: np.ndarray) -> np.ndarray:
    h = x
    for l in range(cfg.num_lstm_layers):
        h = lstm(h, W[pre + f"lstm.{l}.Wx"], W[pre + f"lstm.{l}.Wh"], W[pre + f"lstm.{l}.bias"])
    return h + x


def resnet_block(cfg: EncodecConfig, W: Dict[str, np.ndarray], pre: str, x: np.ndarray, dilations: Sequence[int]) -> np.ndarray:
    h = x
    for bi, d in enumerate(dilations):
        h = conv1d(cfg, elu(h), W[pre + f"block.{2 * bi + 1}.conv.weight"], W[pre + f"block.{2 * bi + 1}.conv.bias"], dilation=d)
    if cfg.use_conv_shortcut:
        return conv1d(cfg, x, W[pre + "shortcut.conv.weight"], W[pre + "shortcut.conv.bias"]) + h
    return x + h


def quantizer_decode(W: Dict[str, np.ndarray], codes: np.ndarray) -> np.ndarray:
    """codes [B, n_q, T] -> [B, T, dim] (EncodecQuantization.swift:117-133)."""
    out = None
    for q in range(codes.shape[1]):
        e = W[f"quantizer.layers.{q}.codebook.embed"].astype(np.float64)[codes[:, q, :]]
        out = e if out is None else out + e
    return out


def decoder(cfg: EncodecConfig, W: Dict[str, np.ndarray], emb: np.ndarray) -> np.ndarray:
    """EncodecDecoder.callAsFunction (Encodec.swift:150-166): [B, T, hidden] -> [B, T*hop, channels]."""
    h = np.asarray(emb, dtype=np.float64)
    for idx, kind, p in decoder_layout(cfg):
        pre = f"decoder.layers.{idx}."
        if kind == "conv":
            h = conv1d(cfg, h, W[pre + "conv.weight"], W[pre + "conv.bias"], dilation=p["dilation"])
        elif kind == "lstm":
            h = lstm_block(cfg, W, pre, h)
        elif kind == "elu":
            h = elu(h)
        elif kind == "convt":
            h = conv_transpose1d(cfg, h, W[pre + "conv.weight"], W[pre + "conv.bias"], p["stride"])
        elif kind == "resnet":
            h = resnet_block(cfg, W, pre, h, p["dilations"])
    return h


def decode_frame(cfg: EncodecConfig, W: Dict[str, np.ndarray], codes: np.ndarray, scale: Optional[np.ndarray] = None) -> np.ndarray:
    """Encodec.decodeFrame (Encodec.swift:294-301)."""
    y = decoder(cfg, W, quantizer_decode(W, codes))
    return y if scale is None else y * np.asarray(scale, dtype=np.float64).reshape(-1, 1, 1)


def linear_overlap_add(frames: Sequence[np.ndarray], hop: int) -> np.ndarray:
    """Encodec.linearOverlapAdd (Encodec.swift:304-356): triangular weights 0.5 - |(t+1)/(L+1) - 0.5|, divide by the weight sum."""
    N, L, C = frames[0].shape
    total = hop * (len(frames) - 1) + frames[-1].shape[1]
    t = (np.arange(L) + 1.0) / (L + 1.0)
    wv = 0.5 - np.abs(t - 0.5)
    out = np.zeros((N, total, C)); sw = np.zeros(total)
    off = 0
    for f in frames:
        fl = f.shape[1]
        out[:, off:off + fl, :] += wv[:fl, None] * f
        sw[off:off + fl] += wv[:fl]
        off += hop
    nz = sw != 0
    out[:, nz, :] /= sw[nz, None]
    return out


def decode(cfg: EncodecConfig, W: Dict[str, np.ndarray], audio_codes: np.ndarray, audio_scales: Optional[Sequence] = None,
           padding_len: Optional[int] = None) -> np.ndarray:
    """Encodec.decode (Encodec.swift:366-402): audio_codes [n_chunks, B, n_q, T] -> [B, samples, channels]."""
    scales = list(audio_scales) if audio_scales is not None else [None] * audio_codes.shape[0]
    if cfg.chunk_length is None:
        assert audio_codes.shape[0] == 1, "Expected one frame"
        y = decode_frame(cfg, W, audio_codes[0], scales[0])
    else:
        y = linear_overlap_add([decode_frame(cfg, W, audio_codes[i], scales[i]) for i in range(audio_codes.shape[0])],
                               cfg.chunk_stride or 1)
    if padding_len is not None and padding_len < y.shape[1]:
        y = y[:, :padding_len, :]
    return y
