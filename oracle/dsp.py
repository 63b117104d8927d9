"""Oracle for the mel front-end.  Test infrastructure only.

Follows, line by line:
  Sources/MLXAudioCore/DSP.swift:15-22     hanningWindow (symmetric)
  Sources/MLXAudioCore/DSP.swift:76-168    melFilters
  Sources/MLXAudioCore/DSP.swift:181-273   stft / computeMelSpectrogram
  Sources/MLXAudioSTT/Streaming/IncrementalMelSpectrogram.swift:43-208
  Sources/MLXAudioSTT/Models/Whisper/WhisperAudio.swift:7-120

Host-side tables (window, filterbank) are computed in float32 exactly as the
Swift ``Float`` code does; the signal path (DFT, power, filterbank product,
log10) is evaluated in float64 and is therefore the "ideal" value the fp32
device path is compared against (tolerance 1e-3 relative, see tests).
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np

f32 = np.float32


def hanning_window(size: int) -> np.ndarray:
    """DSP.swift:15-22 -- symmetric Hann, 0.5*(1-cos(2*pi*n/(N-1))), Float math."""
    n = np.arange(size, dtype=f32)
    denom = f32(size - 1)
    return (f32(0.5) * (f32(1) - np.cos(f32(2) * f32(np.pi) * n / denom, dtype=f32))).astype(f32)


def hamming_window(size: int, periodic: bool = True) -> np.ndarray:
    """DSP.swift:25-42 -- periodic = the first ``size`` points of the (size + 1)-point window."""
    if size <= 0:
        return np.zeros(0)
    if size == 1:
        return np.ones(1)
    eff = size + 1 if periodic else size
    return (0.54 - 0.46 * np.cos(2.0 * np.pi * np.arange(eff) / (eff - 1)))[:size]


def power_to_db(spectrogram, amin: float = 1e-10, top_db=None) -> np.ndarray:
    """DSP.swift:61-73."""
    db = 10.0 * np.log10(np.maximum(np.asarray(spectrogram, dtype=np.float64), amin))
    return db if top_db is None else np.maximum(db, db.max() - top_db)


def periodic_hann_window(size: int) -> np.ndarray:
    """WhisperAudio.swift:42-43 -- periodic Hann, 0.5*(1-cos(2*pi*n/N))."""
    n = np.arange(size, dtype=f32)
    return (f32(0.5) * (f32(1) - np.cos((f32(2) * f32(np.pi) * n) / f32(size), dtype=f32))).astype(f32)


def mel_filters(sample_rate: int, n_fft: int, n_mels: int, f_min: float = 0.0,
                f_max: Optional[float] = None, norm: Optional[str] = "slaney",
                mel_scale: str = "htk") -> np.ndarray:
    """DSP.swift:76-168 -- triangular filterbank [nFreqs, nMels], Float math.

    Note the inclusive upper edge (``<= high``, DSP.swift:149) and that the
    default scale is HTK while Whisper passes ``.slaney`` (WhisperAudio.swift:30).
    """
    f_min = f32(f_min)
    f_max_val = f32(f_max) if f_max is not None else f32(sample_rate) / f32(2.0)
    n_freqs = n_fft // 2 + 1
    all_freqs = (np.arange(n_freqs, dtype=f32) * f32(sample_rate) / f32(n_fft)).astype(f32)

    if mel_scale == "htk":
        def hz_to_mel(freq):
            return f32(2595.0) * f32(np.log10(f32(1.0) + f32(freq) / f32(700.0)))

        def mel_to_hz(mel):
            return f32(700.0) * (f32(np.power(f32(10.0), f32(mel) / f32(2595.0))) - f32(1.0))
    elif mel_scale == "slaney":
        f_sp = f32(200.0) / f32(3.0)
        min_log_hz = f32(1000.0)
        min_log_mel = (min_log_hz - f_min) / f_sp
        log_step = f32(np.log(f32(6.4))) / f32(27.0)

        def hz_to_mel(freq):
            freq = f32(freq)
            if freq < min_log_hz:
                return (freq - f_min) / f_sp
            return min_log_mel + f32(np.log(freq / min_log_hz)) / log_step

        def mel_to_hz(mel):
            mel = f32(mel)
            if mel < min_log_mel:
                return f_min + f_sp * mel
            return min_log_hz * f32(np.exp(log_step * (mel - min_log_mel)))
    else:
        raise ValueError(mel_scale)

    m_min = hz_to_mel(f_min)
    m_max = hz_to_mel(f_max_val)
    m_pts = [f32(m_min + f32(i) * (m_max - m_min) / f32(n_mels + 1)) for i in range(n_mels + 2)]
    f_pts = [f32(mel_to_hz(m)) for m in m_pts]

    fb = np.zeros((n_freqs, n_mels), dtype=f32)
    for i in range(n_freqs):
        fr = all_freqs[i]
        for j in range(n_mels):
            low, center, high = f_pts[j], f_pts[j + 1], f_pts[j + 2]
            if fr >= low and fr < center:
                fb[i, j] = (fr - low) / (center - low)
            elif fr >= center and fr <= high:
                fb[i, j] = (high - fr) / (high - center)
    if norm == "slaney":
        for j in range(n_mels):
            enorm = f32(2.0) / (f_pts[j + 2] - f_pts[j])
            fb[:, j] *= enorm
    return fb


def _frames(signal: np.ndarray, n_frames: int, n_fft: int, hop: int) -> np.ndarray:
    """asStrided(signal, [F, nFft], strides [hop, 1]) (DSP.swift:220, IncrementalMel...:119-124)."""
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    return signal[idx]


def _power_mel_log(signal: np.ndarray, n_frames: int, n_fft: int, hop: int,
                   window: np.ndarray, filters: np.ndarray) -> np.ndarray:
    """frames*window -> rfft -> |.|^2 -> @filters -> max(.,1e-10) -> log10 (float64)."""
    fr = _frames(signal.astype(np.float64), n_frames, n_fft, hop) * window.astype(np.float64)
    spec = np.fft.rfft(fr, axis=1)
    mag = spec.real ** 2 + spec.imag ** 2
    mel = mag @ filters.astype(np.float64)
    return np.log10(np.maximum(mel, 1e-10))


def reflect_pad_core(audio: np.ndarray, padding: int) -> np.ndarray:
    """DSP.swift:194-205 reflect branch (prefix = reversed audio[1:min(p+1,n)], suffix likewise)."""
    n = len(audio)
    prefix = audio[1:min(padding + 1, n)][::-1]
    s0, s1 = max(0, n - padding - 1), max(1, n - 1)
    suffix = audio[s0:s1][::-1]
    return np.concatenate([prefix, audio, suffix])


def compute_mel_spectrogram(audio: np.ndarray, sample_rate: int, n_fft: int, hop: int,
                            n_mels: int) -> np.ndarray:
    """DSP.swift:230-273 -- offline log-mel with GLOBAL max-8 clamp, output [F, nMels]."""
    window = hanning_window(n_fft)
    padded = reflect_pad_core(np.asarray(audio, dtype=f32), n_fft // 2)
    n_frames = 1 + (len(padded) - n_fft) // hop
    filters = mel_filters(sample_rate, n_fft, n_mels, norm="slaney")
    lm = _power_mel_log(padded, n_frames, n_fft, hop, window, filters)
    lm = np.maximum(lm, lm.max() - 8.0)
    return (lm + 4.0) / 4.0


class IncrementalMelSpectrogram:
    """IncrementalMelSpectrogram.swift:18-208 -- overlap-save streaming log-mel.

    ``process`` returns ``None`` where the reference returns ``nil``.
    """

    def __init__(self, sample_rate: int = 16000, n_fft: int = 400, hop_length: int = 160,
                 n_mels: int = 128):
        self.n_fft, self.hop, self.n_mels, self.sr = n_fft, hop_length, n_mels, sample_rate
        self.overlap_size = n_fft - hop_length
        self.window = hanning_window(n_fft)
        self.filters = mel_filters(sample_rate, n_fft, n_mels, norm="slaney")
        self.reset()

    def reset(self) -> None:                      # :203-208
        self.overlap: List[float] = []
        self.is_first = True
        self.running_max = -math.inf
        self.total_frames = 0

    def _emit(self, signal: np.ndarray, n_frames: int) -> np.ndarray:
        lm = _power_mel_log(signal, n_frames, self.n_fft, self.hop, self.window, self.filters)
        self.running_max = max(self.running_max, float(lm.max()))        # :139-140
        lm = np.maximum(lm, self.running_max - 8.0)                       # :142
        self.total_frames += n_frames
        return (lm + 4.0) / 4.0                                           # :143

    def process(self, samples) -> Optional[np.ndarray]:                   # :68-147
        samples = [float(f32(s)) for s in samples]
        if not samples:
            return None
        if self.is_first:
            pad = self.n_fft // 2
            prefix: List[float] = []
            if len(samples) > 1:
                rl = min(pad, len(samples) - 1)
                if rl > 0:
                    prefix = samples[1:rl + 1][::-1]
            if not prefix:
                prefix = [samples[0]] * pad
            elif len(prefix) < pad:
                while len(prefix) < pad:                                   # :88-92
                    needed = pad - len(prefix)
                    prefix = prefix + prefix[:needed]
            signal = prefix + samples
            self.is_first = False
        else:
            signal = self.overlap + samples
        n_frames = max(0, (len(signal) - self.n_fft) // self.hop + 1)
        if n_frames <= 0:
            self.overlap = signal
            return None
        consumed = (n_frames - 1) * self.hop + self.n_fft
        if consumed < len(signal):
            self.overlap = signal[consumed - self.overlap_size:]
        else:
            self.overlap = signal[-self.overlap_size:]
        return self._emit(np.asarray(signal, dtype=f32), n_frames)

    def flush(self) -> Optional[np.ndarray]:                              # :151-200
        if not self.overlap:
            return None
        signal = list(self.overlap)
        if len(signal) < self.n_fft:
            signal += [0.0] * (self.n_fft - len(signal))
        pad = self.n_fft // 2
        n = len(signal)
        rl = min(pad, n - 1)
        suffix = signal[n - 1 - rl:n - 1][::-1]
        signal += suffix
        self.overlap = []
        n_frames = max(0, (len(signal) - self.n_fft) // self.hop + 1)
        if n_frames <= 0:
            return None
        return self._emit(np.asarray(signal, dtype=f32), n_frames)


def whisper_reflect_pad(audio: np.ndarray, pad: int) -> np.ndarray:
    """WhisperAudio.swift:89-112 -- reflect pad with zero-fill fallback."""
    n = len(audio)
    if pad <= 0:
        return audio
    if n <= 1:
        return np.pad(audio, (pad, pad))
    lc = min(pad, n - 1)
    left = audio[1:lc + 1][::-1]
    right = audio[n - 1 - lc:n - 1][::-1]
    pieces = []
    if lc < pad:
        pieces.append(np.zeros(pad - lc, dtype=audio.dtype))
    pieces += [left, audio, right]
    if lc < pad:
        pieces.append(np.zeros(pad - lc, dtype=audio.dtype))
    return np.concatenate(pieces)


WHISPER_SR, WHISPER_NFFT, WHISPER_HOP, WHISPER_CHUNK = 16000, 400, 160, 480000


def whisper_log_mel(audio: np.ndarray, n_mels: int = 80) -> np.ndarray:
    """WhisperAudio.swift:38-79 -- [nMels, nFrames]; periodic Hann, Slaney scale, drops last frame."""
    audio = np.asarray(audio, dtype=f32).reshape(-1)
    window = periodic_hann_window(WHISPER_NFFT)
    padded = whisper_reflect_pad(audio, WHISPER_NFFT // 2)
    ns = len(padded)
    n_frames = 1 + (ns - WHISPER_NFFT) // WHISPER_HOP if ns >= WHISPER_NFFT else 0
    if n_frames <= 0:
        return np.zeros((n_mels, 0))
    filters = mel_filters(WHISPER_SR, WHISPER_NFFT, n_mels, f_min=0.0, f_max=WHISPER_SR / 2.0,
                          norm="slaney", mel_scale="slaney")
    lm = _power_mel_log(padded, n_frames, WHISPER_NFFT, WHISPER_HOP, window, filters)
    lm = lm[:-1]                                                           # :65-67
    if lm.shape[0] == 0:
        return np.zeros((n_mels, 0))
    lm = np.maximum(lm, lm.max() - 8.0)
    return ((lm + 4.0) / 4.0).T


def whisper_encoder_features(audio: np.ndarray, n_mels: int = 80) -> np.ndarray:
    """WhisperAudio.swift:7-13,83-87 -- pad/trim to 480000, -> [1, nFrames, nMels]."""
    audio = np.asarray(audio, dtype=f32).reshape(-1)
    n = len(audio)
    if n > WHISPER_CHUNK:
        audio = audio[:WHISPER_CHUNK]
    elif n < WHISPER_CHUNK:
        audio = np.pad(audio, (0, WHISPER_CHUNK - n))
    return whisper_log_mel(audio, n_mels).T[None]


def synth_audio(n: int, seed: int = 0, sr: int = 16000) -> np.ndarray:
    """BASELINE.md section 3 synthetic audio: 0.5*sin(2*pi*220*t) + 0.1*N(0,1), clipped."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / sr
    x = 0.5 * np.sin(2 * np.pi * 220.0 * t) + 0.1 * rng.standard_normal(n)
    return np.clip(x, -1.0, 1.0).astype(f32)
