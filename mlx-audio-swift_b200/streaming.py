"""Host-side mirror of the streaming STT front end around the incremental mel kernel:

  Sources/MLXAudioSTT/Streaming/StreamingTypes.swift:13-94       DelayPreset, StreamingConfig
  Sources/MLXAudioSTT/Streaming/StreamingEncoder.swift:20-209    StreamingEncoder (window accumulation, overlap, cache)
  Sources/MLXAudioSTT/Streaming/StreamingInferenceSession.swift:993-1068   feedAudio: mel -> windows -> decode cadence

Pure host logic: the encoder is injected (anything with ``n_window_infer`` and ``encode_single_window(frames)``, the two
members StreamingEncoder uses of Qwen3ASRAudioEncoder) and so is the clock, the mel frames come from
``IncrementalMelSpectrogram`` (the CUDA kernel) or any object with the same ``process`` / ``flush`` / ``reset``.  The Qwen3-ASR
model, its tokenizer and the text-merging rules of the session are outside the hot path and not built."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional

import numpy as np

DELAY_PRESETS_MS = {"realtime": 200, "agent": 480, "subtitle": 2400}      # StreamingTypes.swift:13-32


@dataclass
class StreamingConfig:
    """StreamingTypes.swift:37-94 (same defaults).  ``delay_preset``: a preset name or a custom delay in ms."""
    decode_interval_seconds: float = 1.0
    boundary_decode_interval_seconds: float = 0.2
    boundary_boost_seconds: float = 1.0
    encoder_window_overlap_seconds: float = 1.0
    max_cached_windows: int = 60
    delay_preset: object = "agent"
    language: Optional[str] = "English"
    temperature: float = 0.0
    max_tokens_per_pass: int = 512
    min_agreement_passes: int = 2
    boundary_min_agreement_passes: int = 3
    max_decode_windows: int = 1
    finalize_completed_windows: bool = True

    @property
    def delay_ms(self) -> int:
        return DELAY_PRESETS_MS[self.delay_preset] if isinstance(self.delay_preset, str) else int(self.delay_preset)

    def overlap_frames(self, sample_rate: int, hop_length: int = 160) -> int:
        """StreamingInferenceSession.swift:982: max(0, round(overlapSeconds * sampleRate / 160))."""
        return max(0, int(np.floor(self.encoder_window_overlap_seconds * sample_rate / float(hop_length) + 0.5)))


class StreamingEncoder:
    """StreamingEncoder.swift:20-209.  ``encoder.encode_single_window(frames [n, n_mels]) -> [tokens, dim]``."""

    def __init__(self, encoder, max_cached_windows: int = 60, overlap_frames: int = 0):
        self.encoder = encoder
        self.window_size = int(encoder.n_window_infer)
        clamped = max(0, min(int(overlap_frames), max(0, self.window_size - 1)))
        self.window_stride = max(1, self.window_size - clamped)
        self.max_cached_windows = int(max_cached_windows)
        self.reset()

    def reset(self) -> None:                                      # :201-208
        self._cached: List[np.ndarray] = []
        self._new: List[np.ndarray] = []
        self._total = 0
        self._pending: Optional[np.ndarray] = None

    @property
    def _pending_count(self) -> int:
        return 0 if self._pending is None else int(self._pending.shape[0])

    def feed(self, mel_frames) -> int:
        """feed(melFrames:) (:56-97): number of full windows encoded by this call."""
        m = np.asarray(mel_frames)
        self._pending = m if self._pending is None else np.concatenate([self._pending, m], axis=0)
        new = 0
        while self._pending_count >= self.window_size:
            frames = self._pending
            enc = self.encoder.encode_single_window(frames[: self.window_size])
            self._cached.append(enc)
            self._new.append(enc)
            self._total += 1
            new += 1
            self._pending = frames[self.window_stride:] if frames.shape[0] > self.window_stride else None
            if len(self._cached) > self.max_cached_windows:
                self._cached.pop(0)
        return new

    def flush_partial(self) -> int:                               # :101-116
        if self._pending_count == 0:
            return 0
        self._cached.append(self.encoder.encode_single_window(self._pending))
        self._pending = None
        if len(self._cached) > self.max_cached_windows:
            self._cached.pop(0)
        return 1

    def get_cached_encoder_output(self, from_window: Optional[int] = None) -> Optional[np.ndarray]:      # :120-137
        start = 0 if from_window is None else max(0, from_window)
        if start >= len(self._cached):
            return None
        return np.concatenate(self._cached[start:], axis=0)

    def encode_pending(self) -> Optional[np.ndarray]:             # :146-152 (does not consume the pending frames)
        return None if self._pending_count == 0 else self.encoder.encode_single_window(self._pending)

    def get_full_encoder_output(self, from_window: Optional[int] = None) -> Optional[np.ndarray]:        # :157-176
        parts = [p for p in (self.get_cached_encoder_output(from_window), self.encode_pending()) if p is not None]
        return None if not parts else np.concatenate(parts, axis=0)

    @property
    def encoded_window_count(self) -> int:                        # :179-181 (monotonic)
        return self._total

    @property
    def has_pending_frames(self) -> bool:
        return self._pending_count > 0

    def drain_newly_encoded_windows(self) -> List[np.ndarray]:   # :189-193
        out, self._new = self._new, []
        return out

    @property
    def total_cached_tokens(self) -> int:
        return sum(int(w.shape[0]) for w in self._cached)


class StreamingFrontEnd:
    """The part of QwenStreamingInferenceSessionCore.feedAudio (:993-1068) that sits on the hot path: samples -> incremental
    mel -> encoder windows, plus the decision whether a decode pass should be launched now.  ``clock()`` returns seconds."""

    def __init__(self, mel, encoder: StreamingEncoder, config: Optional[StreamingConfig] = None, clock: Optional[Callable[[], float]] = None):
        import time
        self.mel, self.encoder, self.config = mel, encoder, config or StreamingConfig()
        self.clock = clock or time.monotonic
        self.total_samples_fed = 0
        self.is_decoding = False                                  # SessionSharedState.isDecoding: cleared by decode_finished()
        self._last_decode: Optional[float] = None
        self._boost_until: Optional[float] = None
        self._has_new_content = False
        self.last_pass_is_boundary_finalize = False

    def feed_audio(self, samples) -> bool:
        """Returns True when the session would launch a decode pass for this call (and marks one as running)."""
        c = self.config
        samples = np.asarray(samples, dtype=np.float32)
        self.total_samples_fed += int(samples.shape[0])
        frames = self.mel.process(samples)
        if frames is None:
            return False
        new_windows = self.encoder.feed(frames)
        if new_windows > 0 or self.encoder.has_pending_frames:
            self._has_new_content = True
        now = self.clock()
        if new_windows > 0:
            boost = max(0.0, c.boundary_boost_seconds)
            self._boost_until = now + boost if boost > 0 else None
        if self._boost_until is not None and now < self._boost_until:
            interval = min(max(0.05, c.boundary_decode_interval_seconds), max(0.05, c.decode_interval_seconds))
        else:
            self._boost_until = None
            interval = max(0.05, c.decode_interval_seconds)
        if c.finalize_completed_windows and new_windows > 0:
            should = True
        elif self._last_decode is not None:
            should = now - self._last_decode >= interval
        else:
            should = self._has_new_content
        if not (should and self._has_new_content) or self.is_decoding:
            return False
        self.is_decoding = True
        self._has_new_content = False
        self.last_pass_is_boundary_finalize = bool(c.finalize_completed_windows and new_windows > 0)
        if not self.last_pass_is_boundary_finalize:
            self._last_decode = now
        return True

    def min_agreement_passes(self) -> int:
        """launchDecodePassLocked (:1151-1158): the stronger threshold while the boundary boost is active."""
        c = self.config
        if self._boost_until is not None and self.clock() < self._boost_until:
            return max(1, max(c.min_agreement_passes, c.boundary_min_agreement_passes))
        return max(1, c.min_agreement_passes)

    def decode_finished(self) -> None:
        self.is_decoding = False
