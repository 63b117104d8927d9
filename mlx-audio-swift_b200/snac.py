"""Host-side mirror of `SNAC` (Sources/MLXAudioCodecs/SNAC/SNACDecoder.swift:12-204) behind the
AudioCodecModel protocol (Sources/MLXAudioCodecs/AudioCodecModel.swift:4-27), over the C ABI."""
from __future__ import annotations

import ctypes as C
import json
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _ffi


class SNAC:
    def __init__(self, sampling_rate: int = 24000, encoder_dim: int = 48, encoder_rates: Sequence[int] = (2, 4, 8, 8),
                 latent_dim: Optional[int] = None, decoder_dim: int = 1024, decoder_rates: Sequence[int] = (8, 8, 4, 2),
                 attn_window_size: Optional[int] = None, codebook_size: int = 4096, codebook_dim: int = 8,
                 vq_strides: Sequence[int] = (4, 2, 1), noise: bool = True, depthwise: bool = True, *,
                 weights: Dict[str, np.ndarray], device: int = 0):
        cfg = _ffi.SnacConfig()
        cfg.sampling_rate, cfg.encoder_dim = sampling_rate, encoder_dim
        cfg.n_encoder_rates = len(encoder_rates)
        for i, r in enumerate(encoder_rates):
            cfg.encoder_rates[i] = r
        cfg.latent_dim = latent_dim or 0
        cfg.decoder_dim, cfg.n_decoder_rates = decoder_dim, len(decoder_rates)
        for i, r in enumerate(decoder_rates):
            cfg.decoder_rates[i] = r
        cfg.attn_window_size = attn_window_size or 0
        cfg.codebook_size, cfg.codebook_dim, cfg.n_vq_strides = codebook_size, codebook_dim, len(vq_strides)
        for i, r in enumerate(vq_strides):
            cfg.vq_strides[i] = r
        cfg.noise, cfg.depthwise = int(noise), int(depthwise)
        self.sampling_rate, self.vq_strides, self.decoder_rates = sampling_rate, tuple(vq_strides), tuple(decoder_rates)
        self.latent_dim = latent_dim or encoder_dim * 2 ** len(encoder_rates)
        self.n_codebooks = len(vq_strides)
        table, keep = _ffi.make_tensor_table(weights)
        self._h = C.c_void_p()
        _ffi.check(_ffi.lib().b2a_snac_create(device, C.byref(cfg), table, len(weights), C.byref(self._h)))
        del keep
        self.hop_length = int(_ffi.lib().b2a_snac_hop_length(self._h))

    @staticmethod
    def random_init_weights(seed: int = 1234, latent: int = 768, decoder_dim: int = 1024,
                            decoder_rates: Sequence[int] = (8, 8, 4, 2), vq_strides: Sequence[int] = (4, 2, 1),
                            codebook_size: int = 4096, codebook_dim: int = 8) -> Dict[str, np.ndarray]:
        """Random-init snac_24khz-shaped weights (benchmarks): conv v ~ U(+-1/sqrt(fan_in)) as
        Layers.swift:81-86, g = ||v||, zero biases, Snake alpha = 1."""
        rng = np.random.default_rng(seed)
        w: Dict[str, np.ndarray] = {}

        def wn(prefix, shape, fan, bias_n):
            s = (1.0 / fan) ** 0.5
            v = rng.uniform(-s, s, size=shape).astype(np.float32)
            w[prefix + ".weight_v"] = v
            w[prefix + ".weight_g"] = np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True)).astype(np.float32)
            if bias_n:
                w[prefix + ".bias"] = np.zeros(bias_n, dtype=np.float32)

        for i in range(len(vq_strides)):
            q = f"quantizer.quantizers.{i}"
            wn(q + ".in_proj", (codebook_dim, 1, latent), latent, codebook_dim)
            wn(q + ".out_proj", (latent, 1, codebook_dim), codebook_dim, latent)
            w[q + ".codebook.weight"] = rng.standard_normal((codebook_size, codebook_dim)).astype(np.float32)
        p = "decoder.model.layers"
        wn(f"{p}.0", (latent, 7, 1), 7 * latent, latent)
        wn(f"{p}.1", (decoder_dim, 1, latent), latent, decoder_dim)
        li = 2
        for i, s in enumerate(decoder_rates):
            cin, cout = decoder_dim // 2 ** i, decoder_dim // 2 ** (i + 1)
            b = f"{p}.{li}.block.layers"
            w[f"{b}.0.alpha"] = np.ones((1, cin, 1), dtype=np.float32)
            wn(f"{b}.1", (cin, 2 * s, cout), cin * 2 * s, cout)
            wn(f"{b}.2.linear", (cout, 1, cout), cout, 0)
            for j in (3, 4, 5):
                r = f"{b}.{j}.block.layers"
                w[f"{r}.0.alpha"] = np.ones((1, cout, 1), dtype=np.float32)
                wn(f"{r}.1", (cout, 7, 1), cout * 7, cout)
                w[f"{r}.2.alpha"] = np.ones((1, cout, 1), dtype=np.float32)
                wn(f"{r}.3", (cout, 1, cout), cout, cout)
            li += 1
        cf = decoder_dim // 2 ** len(decoder_rates)
        w[f"{p}.{li}.alpha"] = np.ones((1, cf, 1), dtype=np.float32)
        wn(f"{p}.{li + 1}", (1, 7, cf), cf * 7, 1)
        return w

    # -- loading (SNACDecoder.swift:135-189) ------------------------------------------------------
    @classmethod
    def from_config_dict(cls, cfg: dict, weights, device: int = 0) -> "SNAC":
        return cls(cfg["sampling_rate"], cfg["encoder_dim"], cfg["encoder_rates"], cfg.get("latent_dim"),
                   cfg["decoder_dim"], cfg["decoder_rates"], cfg.get("attn_window_size"), cfg["codebook_size"],
                   cfg["codebook_dim"], cfg["vq_strides"], cfg["noise"], cfg["depthwise"], weights=weights, device=device)

    @classmethod
    def from_model_directory(cls, model_dir, device: int = 0) -> "SNAC":
        model_dir = Path(model_dir)
        wpath = model_dir / "model.safetensors"
        if not wpath.exists():
            raise FileNotFoundError(f"Could not find model at {wpath}")   # SNACError.modelNotFound
        from safetensors.numpy import load_file
        cfg = json.loads((model_dir / "config.json").read_text())
        return cls.from_config_dict(cfg, load_file(str(wpath)), device)

    @property
    def stream(self) -> int:
        return int(_ffi.lib().b2a_snac_stream(self._h) or 0)

    # -- AudioCodecModel ------------------------------------------------------------------------
    @property
    def codec_sample_rate(self) -> float:
        return float(self.sampling_rate)

    def decode(self, codes: List[np.ndarray], noise: Optional[List[Optional[np.ndarray]]] = None,
               zero_noise: bool = False, seed: int = 0, out: Optional[np.ndarray] = None) -> np.ndarray:
        """SNAC.decode (:127-131): codes[i] [B, T_i] int -> waveform [B, 1, T*hop] float32.
        `noise[i]` supplies NoiseBlock i's Gaussian draw explicitly."""
        cs = [np.ascontiguousarray(c, dtype=np.int32) for c in codes]
        if len(cs) != self.n_codebooks:
            raise _ffi.AudioGenerationError(_ffi.ERR_INVALID_INPUT, f"expected {self.n_codebooks} code layers")
        B = cs[0].shape[0]
        T = cs[0].shape[1] * self.vq_strides[0]
        for c, s in zip(cs, self.vq_strides):
            if c.shape != (B, T // s):
                raise _ffi.AudioGenerationError(_ffi.ERR_INVALID_INPUT, "code layer shapes do not match vq_strides")
        cp = (C.c_void_p * len(cs))(*[c.ctypes.data for c in cs])
        nz, np_ = None, None
        if noise is not None:
            nz = [None if n is None else np.ascontiguousarray(n, dtype=np.float32) for n in noise]
            np_ = (C.c_void_p * len(self.decoder_rates))(*[None if n is None else n.ctypes.data for n in nz])
        wave = out if out is not None else np.empty((B, 1, T * self.hop_length), dtype=np.float32)   # `out`: a caller-owned (e.g. pinned) buffer
        assert wave.shape == (B, 1, T * self.hop_length) and wave.dtype == np.float32 and wave.flags["C_CONTIGUOUS"]
        _ffi.check(_ffi.lib().b2a_snac_decode(self._h, cp, B, T, np_, int(zero_noise), seed, _ffi.ptr(wave)))
        return wave

    def decode_audio(self, codes):            # AudioDecoderModel.decodeAudio (:199)
        return self.decode(codes)

    def decode_dev(self, d_codes, d_wave, zero_noise: bool = False, seed: int = 0, stream: int = 0) -> None:
        """Device-resident decode: torch CUDA int32 tensors [B, T_i] -> d_wave [B, 1, T*hop]."""
        B = d_codes[0].shape[0]
        T = d_codes[0].shape[1] * self.vq_strides[0]
        cp = (C.c_void_p * len(d_codes))(*[c.data_ptr() for c in d_codes])
        _ffi.check(_ffi.lib().b2a_snac_decode_dev(self._h, cp, B, T, None, int(zero_noise), seed, _ffi.ptr(d_wave),
                                                  C.c_void_p(stream)))

    def quantize(self, z: np.ndarray):
        """ResidualVectorQuantize.callAsFunction (VQ.swift:150-163): z [B, D, T] -> (z_q, [codes_i])."""
        z = np.ascontiguousarray(z, dtype=np.float32)
        B, D, T = z.shape
        codes = [np.empty((B, T // s), dtype=np.int32) for s in self.vq_strides]
        cp = (C.c_void_p * len(codes))(*[c.ctypes.data for c in codes])
        zq = np.empty_like(z)
        _ffi.check(_ffi.lib().b2a_snac_quantize(self._h, _ffi.ptr(z), B, T, cp, _ffi.ptr(zq)))
        return zq, codes

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                _ffi.lib().b2a_snac_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:   # interpreter shutdown: ctypes globals may already be gone
            pass
