"""Host-side mirror of `WhisperModel` (Sources/MLXAudioSTT/Models/Whisper/WhisperModel.swift:7-309) behind
STTGenerationModel (Sources/MLXAudioSTT/Generation.swift:52-64), over the C ABI.  Tokenisation / detokenisation stay
with the host (`WhisperTokenizer`); `generate` returns token ids plus the STTOutput statistics."""
from __future__ import annotations

import ctypes as C
import time
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _ffi

# multilingual vocabulary ids (WhisperTokenizer.swift)
EOT, SOT, TRANSLATE, TRANSCRIBE, NO_TIMESTAMPS, TIMESTAMP_BEGIN = 50257, 50258, 50358, 50359, 50363, 50364
LANG_EN = 50259
CHUNK_SAMPLES = 480000


@dataclass
class STTGenerateParameters:
    """Sources/MLXAudioSTT/Generation.swift:3-50 as used by Whisper (defaultGenerationParameters, WhisperModel.swift:21-34)."""
    max_tokens: int = 448 - 16
    temperature: float = 0.0
    language_id: Optional[int] = LANG_EN      # already resolved to its token id; None = let the model decide
    task: str = "transcribe"
    begin_suppress_tokens: Sequence[int] = (EOT,)
    suppress_tokens: Sequence[int] = ()
    seed: int = 0                             # temperature > 0: draws are a pure function of (seed, clip, step)
    mask_eot: bool = False                    # benchmark only (b2a_stt_set_bench_flags, include/b200audio_internal.h)


@dataclass
class STTOutput:
    """Sources/MLXAudioSTT/Models/GLMASR/STTOutput.swift:80-133 (text is left to the host tokenizer)."""
    tokens: List[List[int]]
    prompt_tokens: int
    generation_tokens: int
    total_tokens: int
    prompt_tps: float
    generation_tps: float
    total_time: float
    encode_time: float = 0.0
    decode_time: float = 0.0
    segments: Optional[list] = None


def build_prompt_tokens(language_id: Optional[int] = LANG_EN, task: str = "transcribe", multilingual: bool = True) -> List[int]:
    """WhisperTokenizer.buildPromptTokens (WhisperTokenizer.swift:98-113) on resolved ids."""
    toks = [SOT]
    if multilingual:
        if language_id is not None:
            toks.append(language_id)
        toks.append(TRANSLATE if task.lower() == "translate" else TRANSCRIBE)
    toks.append(NO_TIMESTAMPS)
    return toks


class WhisperModel:
    sample_rate = 16000

    @staticmethod
    def _c_config(config: dict, max_batch: int) -> _ffi.WhisperConfig:
        d = config.get("d_model", 384)
        return _ffi.WhisperConfig(config.get("vocab_size", 51865), config.get("num_mel_bins", 80), d,
                                  config.get("encoder_layers", 4), config.get("encoder_attention_heads", 6),
                                  config.get("encoder_ffn_dim", 4 * d), config.get("max_source_positions", 1500),
                                  config.get("decoder_layers", 4), config.get("decoder_attention_heads", 6),
                                  config.get("decoder_ffn_dim", 4 * d), config.get("max_target_positions", 448), max_batch)

    def __init__(self, config: dict, weights: Dict, device: int = 0, max_batch: int = 16):
        self.config = config
        # sanitizeHuggingFace (WhisperModel.swift:335-365): drop the tied proj_out, accept keys without "model."
        w = {}
        for k, v in weights.items():
            if k in ("proj_out.weight", "model.proj_out.weight"):
                continue
            if not k.startswith("model.") and (k.startswith("encoder.") or k.startswith("decoder.")):
                k = "model." + k
            w[k] = v
        c = self._c_config(config, max_batch)
        table, keep = _ffi.make_tensor_table(w)
        self._h = C.c_void_p()
        _ffi.check(_ffi.lib().b2a_stt_create(device, C.byref(c), table, len(w), C.byref(self._h)))
        del keep

    @classmethod
    def from_model_directory(cls, model_dir, device: int = 0, max_batch: int = 16) -> "WhisperModel":
        """fromDirectory + sanitize (WhisperModel.swift:286-333): config.json + *.safetensors in either the HF (`openai/whisper-*`) or
        the mlx-whisper (`mlx-community/whisper-*`) key layout; the remap, the conv re-layout and the sinusoid synthesis run in the
        library (b2a_weights_sanitize_whisper)."""
        import json
        from pathlib import Path
        from .loading import Weights
        config = json.loads((Path(model_dir) / "config.json").read_text())
        quant = config.get("quantization")         # quantised checkpoints: every Linear + decoder.embed_tokens (WhisperModel.swift:499-511)
        if "n_audio_state" in config:            # mlx-whisper / OpenAI dims (WhisperConfig.swift:94-131 accepts both spellings)
            config = dict(vocab_size=config["n_vocab"], num_mel_bins=config["n_mels"], d_model=config["n_audio_state"],
                          encoder_layers=config["n_audio_layer"], encoder_attention_heads=config["n_audio_head"],
                          encoder_ffn_dim=4 * config["n_audio_state"], max_source_positions=config["n_audio_ctx"],
                          decoder_layers=config["n_text_layer"], decoder_attention_heads=config["n_text_head"],
                          decoder_ffn_dim=4 * config["n_text_state"], max_target_positions=config["n_text_ctx"])
        w = Weights(model_dir)
        w.sanitize_whisper()
        if isinstance(quant, dict):
            w.dequantize(int(quant["group_size"]), int(quant["bits"]))
        tensors = w.tensors()
        w.close()
        return cls(config, tensors, device=device, max_batch=max_batch)

    @classmethod
    def random_init(cls, config: dict, device: int = 0, max_batch: int = 16, std: float = 0.05, seed: int = 1234):
        self = cls.__new__(cls)
        self.config = config
        c = cls._c_config(config, max_batch)
        self._h = C.c_void_p()
        _ffi.check(_ffi.lib().b2a_stt_create_random(device, C.byref(c), std, seed, C.byref(self._h)))
        return self

    @property
    def stream(self) -> int:
        return int(_ffi.lib().b2a_stt_stream(self._h) or 0)

    @property
    def default_generation_parameters(self) -> STTGenerateParameters:
        return STTGenerateParameters(max_tokens=self.config.get("max_target_positions", 448) - 16)

    @staticmethod
    def _clips(audio) -> np.ndarray:
        x = np.ascontiguousarray(audio, dtype=np.float32)
        if x.ndim == 1:
            x = x[None]
        return x

    def encode(self, audio) -> np.ndarray:
        """model.encoder(WhisperAudio.encoderFeatures(audio)) -> [B, 1500, d_model]."""
        x = self._clips(audio)
        out = np.empty((x.shape[0], 1500, self.config.get("d_model", 384)), dtype=np.float32)
        _ffi.check(_ffi.lib().b2a_stt_encode(self._h, _ffi.ptr(x), x.shape[0], x.shape[1], _ffi.ptr(out)))
        return out

    def decoder_logits(self, tokens) -> np.ndarray:
        """Teacher-forced decoder pass against the last `encode`: tokens [B, T] -> logits [B, T, vocab]."""
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        out = np.empty((t.shape[0], t.shape[1], self.config.get("vocab_size", 51865)), dtype=np.float32)
        _ffi.check(_ffi.lib().b2a_stt_decoder_logits(self._h, _ffi.ptr(t), t.shape[0], t.shape[1], _ffi.ptr(out)))
        return out

    def _params(self, p: STTGenerateParameters):
        prompt = np.asarray(build_prompt_tokens(p.language_id, p.task), dtype=np.int32)
        bs = np.asarray(list(p.begin_suppress_tokens), dtype=np.int32)
        su = np.asarray(list(p.suppress_tokens), dtype=np.int32)
        sp = _ffi.SttParams(p.max_tokens, p.temperature, prompt.ctypes.data, len(prompt), bs.ctypes.data if len(bs) else None,
                            len(bs), su.ctypes.data if len(su) else None, len(su), TIMESTAMP_BEGIN, EOT, int(p.seed))
        _ffi.check(_ffi.lib().b2a_stt_set_bench_flags(self._h, int(p.mask_eot)))
        return sp, (prompt, bs, su)

    def generate(self, audio, generation_parameters: Optional[STTGenerateParameters] = None) -> STTOutput:
        """generate(audio:generationParameters:) (WhisperModel.swift:36-93) for a batch of <=30 s clips [B, n]."""
        p = generation_parameters or self.default_generation_parameters
        x = self._clips(audio)
        B = x.shape[0]
        sp, keep = self._params(p)
        toks = np.zeros((B, p.max_tokens), dtype=np.int32)
        ntok = np.zeros(B, dtype=np.int32)
        info = _ffi.SttInfo()
        _ffi.check(_ffi.lib().b2a_stt_transcribe(self._h, _ffi.ptr(x), B, x.shape[1], C.byref(sp), _ffi.ptr(toks), _ffi.ptr(ntok),
                                                 C.byref(info)))
        del keep
        tt = max(info.total_time, 1e-9)
        return STTOutput([toks[b, :ntok[b]].tolist() for b in range(B)], info.prompt_tokens, info.generation_tokens,
                         info.prompt_tokens + info.generation_tokens, info.prompt_tokens / tt, info.generation_tokens / tt,
                         info.total_time, info.encode_time, info.decode_time)

    def generate_dev(self, d_pcm, generation_parameters: STTGenerateParameters, tokens_host, n_tokens_host) -> STTOutput:
        """Device-resident audio (torch CUDA tensor [B, n]); token ids land in the caller's host arrays."""
        sp, keep = self._params(generation_parameters)
        B, n = d_pcm.shape
        info = _ffi.SttInfo()
        _ffi.check(_ffi.lib().b2a_stt_transcribe_dev(self._h, _ffi.ptr(d_pcm), B, n, C.byref(sp), _ffi.ptr(tokens_host),
                                                     _ffi.ptr(n_tokens_host), C.byref(info)))
        del keep
        tt = max(info.total_time, 1e-9)
        return STTOutput([], info.prompt_tokens, info.generation_tokens, info.prompt_tokens + info.generation_tokens,
                         info.prompt_tokens / tt, info.generation_tokens / tt, info.total_time, info.encode_time, info.decode_time)

    def generate_long(self, audio, generation_parameters: Optional[STTGenerateParameters] = None) -> STTOutput:
        """generate(audio:) for ONE mono signal of any length (WhisperModel.swift:95-182): consecutive 30 s windows
        (chunkAudioFor30sWindows :165-182), transcribed as a batch; `segments` carries (start, end) seconds per window like the
        reference's allSegments; `tokens` is the per-window id lists (detokenise and join with " " on the host)."""
        p = generation_parameters or self.default_generation_parameters
        x = np.ascontiguousarray(audio, dtype=np.float32)
        if x.ndim > 1:
            x = np.ascontiguousarray(x.mean(axis=-1), dtype=np.float32)          # mono = audio.mean(axis: -1) (:98)
        n = x.shape[0]
        max_chunks = max(1, -(-n // 480000))
        sp, keep = self._params(p)
        toks = np.zeros((max_chunks, p.max_tokens), dtype=np.int32)
        ntok = np.zeros(max_chunks, dtype=np.int32)
        offs = np.zeros(max_chunks, dtype=np.float32)
        nch = C.c_int32(0)
        info = _ffi.SttInfo()
        _ffi.check(_ffi.lib().b2a_stt_transcribe_long(self._h, _ffi.ptr(x), n, C.byref(sp), max_chunks, _ffi.ptr(toks), _ffi.ptr(ntok),
                                                      _ffi.ptr(offs), C.byref(nch), C.byref(info)))
        del keep
        k = int(nch.value)
        tt = max(info.total_time, 1e-9)
        segs = [{"start": float(offs[i]), "end": float(offs[i]) + min(480000, n - i * 480000) / 16000.0} for i in range(k)]
        return STTOutput([toks[i, :ntok[i]].tolist() for i in range(k)], info.prompt_tokens, info.generation_tokens,
                         info.prompt_tokens + info.generation_tokens, info.prompt_tokens / tt, info.generation_tokens / tt,
                         info.total_time, info.encode_time, info.decode_time, segs)

    def generate_stream(self, audio, generation_parameters: Optional[STTGenerateParameters] = None):
        """generateStream(audio:generationParameters:) (WhisperModel.swift:92-156) on token ids: ('token', id) for every generated token,
        window by window (the reference yields the detokenised deltas), then ('result', STTOutput).  The windows are transcribed as one
        batch, so the token events arrive after the call returns, like LlamaTTSModel.generate_stream."""
        out = self.generate_long(audio, generation_parameters)
        for window in out.tokens:
            for tok in window:
                yield ("token", tok)
        yield ("result", out)

    def cancel(self) -> None:
        _ffi.check(_ffi.lib().b2a_stt_cancel(self._h))

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                _ffi.lib().b2a_stt_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------ streaming session (row N3)
@dataclass
class StreamingConfig:
    """StreamingConfig (StreamingTypes.swift:36-92), the fields the generic session core reads."""
    decode_interval_seconds: float = 1.0
    window_seconds: float = 8.0
    encoder_window_overlap_seconds: float = 1.0
    delay_ms: int = 480                      # DelayPreset.agent; realtime 200, subtitle 2400
    min_agreement_passes: int = 2
    max_tokens_per_pass: int = 512
    sample_rate: int = 16000

    def _c(self) -> "_ffi.SttStreamConfig":
        return _ffi.SttStreamConfig(self.decode_interval_seconds, self.window_seconds, self.encoder_window_overlap_seconds, self.delay_ms,
                                    self.min_agreement_passes, self.max_tokens_per_pass, self.sample_rate)


@dataclass
class StreamingUpdate:
    kind: str                                # "none" | "partial" | "final_window" | "ended"
    promoted: int
    completed: List[List[int]]
    confirmed: List[int]
    provisional: List[int]
    total_audio_seconds: float
    pass_encode_time: float
    pass_decode_time: float


class StreamingInferenceSession:
    """StreamingInferenceSession(model:config:) for `any STTGenerationModel` (StreamingInferenceSession.swift:162, core :589-950) at the
    token level: `feed_audio(samples, now)` / `stop(now)` return what the reference would turn into TranscriptionEvents.  `model` is a
    WhisperModel, or `decoder(audio, prefix) -> continuation` for a host-side model."""

    _KINDS = ("none", "partial", "final_window", "ended")

    def __init__(self, model=None, config: Optional[StreamingConfig] = None, generation_parameters: Optional[STTGenerateParameters] = None,
                 decoder=None):
        self._h = C.c_void_p()
        cfg = (config or StreamingConfig())._c()
        self._keep = None
        if decoder is not None:
            def _cb(user, pcm, n, prefix, n_prefix, out, cap, n_out):
                try:
                    toks = list(decoder(np.ctypeslib.as_array(pcm, shape=(n,)).copy() if n else np.zeros(0, np.float32),
                                        [int(prefix[i]) for i in range(n_prefix)]))[:cap]
                    for i, t in enumerate(toks):
                        out[i] = int(t)
                    n_out[0] = len(toks)
                    return 0
                except Exception:      # the C side reports generationFailed
                    return 1
            self._keep = _ffi.STT_DECODE_CB(_cb)
            _ffi.check(_ffi.lib().b2a_stt_session_create_with_decoder(self._keep, None, C.byref(cfg), C.byref(self._h)))
        else:
            p = generation_parameters or model.default_generation_parameters
            sp, keep = model._params(p)
            _ffi.check(_ffi.lib().b2a_stt_session_create(model._h, C.byref(sp), C.byref(cfg), C.byref(self._h)))
            self._model = model            # the session borrows the model handle

    def _tokens(self, which: int, window: int = 0) -> List[int]:
        n = C.c_int32(0)
        _ffi.check(_ffi.lib().b2a_stt_session_tokens(self._h, which, window, None, 0, C.byref(n)))
        buf = np.zeros(max(1, n.value), dtype=np.int32)
        _ffi.check(_ffi.lib().b2a_stt_session_tokens(self._h, which, window, _ffi.ptr(buf), len(buf), C.byref(n)))
        return buf[:n.value].tolist()

    def _update(self, u) -> StreamingUpdate:
        return StreamingUpdate(self._KINDS[u.kind], u.promoted, [self._tokens(0, w) for w in range(u.completed_windows)], self._tokens(1),
                               self._tokens(2), u.total_audio_s, u.pass_encode_time, u.pass_decode_time)

    def feed_audio(self, samples, now: Optional[float] = None) -> StreamingUpdate:
        x = np.ascontiguousarray(samples, dtype=np.float32).reshape(-1)
        u = _ffi.SttStreamUpdate()
        _ffi.check(_ffi.lib().b2a_stt_session_feed(self._h, _ffi.ptr(x) if len(x) else None, len(x), time.monotonic() if now is None else float(now), C.byref(u)))
        return self._update(u)

    def stop(self, now: Optional[float] = None) -> StreamingUpdate:
        u = _ffi.SttStreamUpdate()
        _ffi.check(_ffi.lib().b2a_stt_session_stop(self._h, time.monotonic() if now is None else float(now), C.byref(u)))
        return self._update(u)

    def __del__(self):
        if getattr(self, "_h", None):
            _ffi.lib().b2a_stt_session_destroy(self._h)
            self._h = None
