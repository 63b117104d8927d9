"""Host-side mirror of the reference's mel front-end API over the C ABI.

  hanning_window / mel_filters       Sources/MLXAudioCore/DSP.swift:15-22, 76-168
  IncrementalMelSpectrogram          Sources/MLXAudioSTT/Streaming/IncrementalMelSpectrogram.swift:18-208
  compute_mel_spectrogram            Sources/MLXAudioCore/DSP.swift:230-273
  WhisperAudio.encoder_features      Sources/MLXAudioSTT/Models/Whisper/WhisperAudio.swift:83-87
Same names, argument meaning and nil/None behaviour as the Swift API; arrays are numpy float32."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _ffi


def hanning_window(size: int, periodic: bool = False) -> np.ndarray:
    out = np.empty(size, dtype=np.float32)
    _ffi.check(_ffi.lib().b2a_hanning_window(size, int(periodic), _ffi.ptr(out)))
    return out


def hamming_window(size: int, periodic: bool = True) -> np.ndarray:
    """hammingWindow(size:periodic:) (Sources/MLXAudioCore/DSP.swift:25-42)."""
    out = np.empty(max(size, 0), dtype=np.float32)
    _ffi.check(_ffi.lib().b2a_hamming_window(size, int(periodic), _ffi.ptr(out) if size > 0 else None))
    return out


def power_to_db(spectrogram, amin: float = 1e-10, top_db: Optional[float] = None) -> np.ndarray:
    """powerToDB(_:amin:topDB:) (DSP.swift:61-73)."""
    x = np.ascontiguousarray(spectrogram, dtype=np.float32)
    out = np.empty_like(x)
    _ffi.check(_ffi.lib().b2a_power_to_db(_ffi.ptr(x), x.size, amin, -1.0 if top_db is None else float(top_db), _ffi.ptr(out)))
    return out


def mel_filters(sample_rate: int, n_fft: int, n_mels: int, f_min: float = 0.0, f_max: Optional[float] = None,
                norm: Optional[str] = "slaney", mel_scale: str = "htk") -> np.ndarray:
    out = np.empty((n_fft // 2 + 1, n_mels), dtype=np.float32)
    _ffi.check(_ffi.lib().b2a_mel_filters(sample_rate, n_fft, n_mels, f_min, -1.0 if f_max is None else f_max,
                                          int(norm == "slaney"), {"htk": 0, "slaney": 1}[mel_scale], _ffi.ptr(out)))
    return out


class IncrementalMelSpectrogram:
    """init(sampleRate:nFft:hopLength:nMels:) / process(samples:) / flush() / reset() / totalFrames."""

    def __init__(self, sample_rate: int = 16000, n_fft: int = 400, hop_length: int = 160, n_mels: int = 128,
                 device: int = 0):
        self.n_mels = n_mels
        self._h = C.c_void_p()
        _ffi.check(_ffi.lib().b2a_mel_create(device, sample_rate, n_fft, hop_length, n_mels, C.byref(self._h)))

    def process(self, samples) -> Optional[np.ndarray]:
        x = np.ascontiguousarray(samples, dtype=np.float32).reshape(-1)
        cap = int(_ffi.lib().b2a_mel_max_frames(self._h, len(x)))
        out = np.empty((cap, self.n_mels), dtype=np.float32)
        n = C.c_int64(0)
        _ffi.check(_ffi.lib().b2a_mel_process(self._h, _ffi.ptr(x), len(x), _ffi.ptr(out), cap, C.byref(n)))
        return out[:n.value].copy() if n.value > 0 else None

    def flush(self) -> Optional[np.ndarray]:
        cap = int(_ffi.lib().b2a_mel_max_frames(self._h, 0))
        out = np.empty((cap, self.n_mels), dtype=np.float32)
        n = C.c_int64(0)
        _ffi.check(_ffi.lib().b2a_mel_flush(self._h, _ffi.ptr(out), cap, C.byref(n)))
        return out[:n.value].copy() if n.value > 0 else None

    def reset(self) -> None:
        _ffi.check(_ffi.lib().b2a_mel_reset(self._h))

    @property
    def total_frames(self) -> int:
        return int(_ffi.lib().b2a_mel_total_frames(self._h))

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                _ffi.lib().b2a_mel_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:   # interpreter shutdown: ctypes globals may already be gone
            pass


class LogMel:
    """Batched offline log-mel. kind 'core' = computeMelSpectrogram, 'whisper' = WhisperAudio.encoderFeatures."""

    def __init__(self, kind: str = "whisper", sample_rate: int = 16000, n_fft: int = 400, hop_length: int = 160,
                 n_mels: int = 80, device: int = 0):
        self.n_mels, self.kind = n_mels, kind
        self._h = C.c_void_p()
        _ffi.check(_ffi.lib().b2a_logmel_create(device, {"core": 0, "whisper": 1}[kind], sample_rate, n_fft, hop_length,
                                                n_mels, C.byref(self._h)))

    def frames(self, n_samples: int) -> int:
        return int(_ffi.lib().b2a_logmel_frames(self._h, n_samples))

    def __call__(self, pcm) -> np.ndarray:
        """pcm [B, n] float32 (host) -> [B, F, n_mels]."""
        x = np.ascontiguousarray(pcm, dtype=np.float32)
        if x.ndim == 1:
            x = x[None]
        B, n = x.shape
        out = np.empty((B, self.frames(n), self.n_mels), dtype=np.float32)
        _ffi.check(_ffi.lib().b2a_logmel_compute(self._h, _ffi.ptr(x), B, n, _ffi.ptr(out)))
        return out

    def compute_dev(self, d_pcm, d_out, stream: int = 0) -> None:
        """Device-resident variant: torch CUDA tensors [B, n] -> [B, F, n_mels] (written in place)."""
        B, n = d_pcm.shape
        _ffi.check(_ffi.lib().b2a_logmel_compute_dev(self._h, _ffi.ptr(d_pcm), B, n, _ffi.ptr(d_out), C.c_void_p(stream)))

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                _ffi.lib().b2a_logmel_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:   # interpreter shutdown: ctypes globals may already be gone
            pass


def compute_mel_spectrogram(audio, sample_rate: int, n_fft: int, hop_length: int, n_mels: int, device: int = 0):
    """DSP.swift:230-273 -> [F, n_mels]."""
    return LogMel("core", sample_rate, n_fft, hop_length, n_mels, device)(np.asarray(audio, dtype=np.float32))[0]


def whisper_encoder_features(audio, n_mels: int = 80, device: int = 0) -> np.ndarray:
    """WhisperAudio.encoderFeatures -> [1, 3000, n_mels]."""
    return LogMel("whisper", 16000, 400, 160, n_mels, device)(np.asarray(audio, dtype=np.float32).reshape(1, -1))
