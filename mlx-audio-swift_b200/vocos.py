"""Host-side mirror of `Vocos` (Sources/MLXAudioCodecs/Vocos/Vocos.swift:284-322) behind AudioDecoderModel, over the C ABI."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import _ffi


class Vocos:
    """Vocos(backbone: VocosBackbone(inputChannels:dim:intermediateDim:numLayers:...), head: ISTFTHead(dim:nFft:hopLength:))."""

    def __init__(self, input_channels: int, dim: int, intermediate_dim: int, num_layers: int, n_fft: int, hop_length: int,
                 input_kernel_size: int = 7, dw_kernel_size: int = 7, adanorm_num_embeddings: Optional[int] = None, *,
                 weights: Dict[str, np.ndarray], device: int = 0):
        cfg = _ffi.VocosConfig(input_channels, dim, intermediate_dim, num_layers, n_fft, hop_length, input_kernel_size,
                               dw_kernel_size, adanorm_num_embeddings or 0)
        self.input_channels, self.hop_length, self.n_fft = input_channels, hop_length, n_fft
        self.adanorm_num_embeddings = adanorm_num_embeddings or 0
        table, keep = _ffi.make_tensor_table(weights)
        self._h = C.c_void_p()
        _ffi.check(_ffi.lib().b2a_vocos_create(device, C.byref(cfg), table, len(weights), C.byref(self._h)))
        del keep

    @property
    def codec_sample_rate(self):          # AudioDecoderModel.codecSampleRate is nil for Vocos (:316)
        return None

    @property
    def stream(self) -> int:
        return int(_ffi.lib().b2a_vocos_stream(self._h) or 0)

    def decode(self, features, bandwidth_id=None) -> np.ndarray:
        """decode(_ features:bandwidthId:) (:302-306): [B, L, C] (or [B, C, L], transposed like VocosBackbone.swift:170-175) -> [B, (L-1)*hop].
        `bandwidth_id` [B, adanorm_num_embeddings]: the conditioning rows of an AdaLayerNorm model (Vocos.swift:17-47)."""
        x = np.ascontiguousarray(features, dtype=np.float32)
        if x.ndim == 2:
            x = x[None]
        if x.shape[-1] != self.input_channels:
            x = np.ascontiguousarray(x.transpose(0, 2, 1))
        B, L, _ = x.shape
        n = int(_ffi.lib().b2a_vocos_output_length(self._h, L))
        out = np.empty((B, n), dtype=np.float32)
        if bandwidth_id is not None:
            c = np.ascontiguousarray(bandwidth_id, dtype=np.float32).reshape(B, -1)
            _ffi.check(_ffi.lib().b2a_vocos_decode_cond(self._h, _ffi.ptr(x), _ffi.ptr(c), B, L, _ffi.ptr(out)))
        else:
            _ffi.check(_ffi.lib().b2a_vocos_decode(self._h, _ffi.ptr(x), B, L, _ffi.ptr(out)))
        return out

    def decode_audio(self, features) -> np.ndarray:
        return self.decode(features)

    def decode_dev(self, d_features, d_wave, stream: int = 0) -> None:
        B, L, _ = d_features.shape
        _ffi.check(_ffi.lib().b2a_vocos_decode_dev(self._h, _ffi.ptr(d_features), B, L, _ffi.ptr(d_wave), C.c_void_p(stream)))

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                _ffi.lib().b2a_vocos_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass
