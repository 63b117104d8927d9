"""Host-side mirror of Qwen3-TTS's autoregressive half over the C ABI:
`Qwen3TTSTalkerForConditionalGeneration` + `Qwen3TTSCodePredictor` + the frame loop / `sampleToken` of `Qwen3TTSModel.generate`
(Sources/MLXAudioTTS/Models/Qwen3TTS/{Qwen3TTSTalker,Qwen3TTSCodePredictor,Qwen3TTS}.swift).  Every number comes from the library
(`b2a_qwen3_talker_*`); this file only composes the prompt rows the way `prepareGenerationInputs` does (Qwen3TTS.swift:883-999) and
chains the speech-tokenizer decoder for audio.  Tokenisation stays with the host tokenizer: the entry points take token ids."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Callable, Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np

from . import _ffi
from .llama_tts import AudioGenerationInfo


@dataclass
class Qwen3CodePredictorConfig:
    """Qwen3TTSConfig.swift:45-63."""
    vocab_size: int = 2048
    hidden_size: int = 1024
    intermediate_size: int = 3072
    num_hidden_layers: int = 5
    num_attention_heads: int = 16
    num_key_value_heads: int = 8
    head_dim: int = 128
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1_000_000.0
    num_code_groups: int = 16


@dataclass
class Qwen3TalkerConfig:
    """Qwen3TTSConfig.swift:268-292."""
    vocab_size: int = 3072
    hidden_size: int = 1024
    intermediate_size: int = 3072
    num_hidden_layers: int = 28
    num_attention_heads: int = 16
    num_key_value_heads: int = 8
    head_dim: int = 128
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1_000_000.0
    num_code_groups: int = 16
    text_hidden_size: int = 2048
    text_vocab_size: int = 151936
    codec_eos_token_id: int = 2150
    code_predictor: Qwen3CodePredictorConfig = field(default_factory=Qwen3CodePredictorConfig)
    # special ids of the codec prefix (Qwen3TTSConfig.swift:294-300)
    codec_think_id: int = 2154
    codec_nothink_id: int = 2155
    codec_think_bos_id: int = 2156
    codec_think_eos_id: int = 2157
    codec_pad_id: int = 2148
    codec_bos_id: int = 2149

    def _c(self, max_batch: int, max_context: int) -> _ffi.Qwen3TalkerConfig:
        cp = self.code_predictor
        return _ffi.Qwen3TalkerConfig(self.vocab_size, self.hidden_size, self.intermediate_size, self.num_hidden_layers,
                                      self.num_attention_heads, self.num_key_value_heads, self.head_dim, self.rms_norm_eps, self.rope_theta,
                                      self.num_code_groups, self.text_hidden_size, self.text_vocab_size, self.codec_eos_token_id,
                                      cp.vocab_size, cp.hidden_size, cp.intermediate_size, cp.num_hidden_layers, cp.num_attention_heads,
                                      cp.num_key_value_heads, cp.head_dim, cp.rms_norm_eps, cp.rope_theta, max_batch, max_context)


@dataclass
class Qwen3GenerateParameters:
    """Qwen3TTSModel.defaultGenerationParameters + sampleToken's arguments (Qwen3TTS.swift:360-385, 1003-1118)."""
    max_tokens: int = 4096
    temperature: float = 0.9
    top_p: float = 1.0
    top_k: int = 50
    min_p: float = 0.0
    repetition_penalty: float = 1.05
    seed: int = 0
    mask_eos: bool = False          # benchmark only (b2a_qwen3_talker_set_bench_flags, include/b200audio_internal.h)

    def _c(self) -> _ffi.Qwen3GenParams:
        return _ffi.Qwen3GenParams(self.max_tokens, self.temperature, self.top_p, self.top_k, self.min_p, self.repetition_penalty, self.seed)


FRAME_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32))


class Qwen3TTSTalker:
    """Talker + code predictor behind one handle.  `weights`: the reference's keys after sanitize strips "talker."."""

    def __init__(self, config: Qwen3TalkerConfig, weights: Dict, device: int = 0, max_batch: int = 8, max_context: int = 2048):
        self.config = config
        c = config._c(max_batch, max_context)
        table, keep = _ffi.make_tensor_table(weights)
        self._h = C.c_void_p()
        _ffi.check(_ffi.lib().b2a_qwen3_talker_create(device, C.byref(c), table, len(weights), C.byref(self._h)))
        del keep

    @classmethod
    def random_init(cls, config: Qwen3TalkerConfig, device: int = 0, max_batch: int = 8, max_context: int = 2048, std: float = 0.02,
                    seed: int = 1234) -> "Qwen3TTSTalker":
        self = cls.__new__(cls)
        self.config = config
        c = config._c(max_batch, max_context)
        self._h = C.c_void_p()
        _ffi.check(_ffi.lib().b2a_qwen3_talker_create_random(device, C.byref(c), std, seed, C.byref(self._h)))
        return self

    @classmethod
    def from_model_directory(cls, model_dir, device: int = 0, max_batch: int = 8, max_context: int = 2048) -> "Qwen3TTSTalker":
        """The talker half of Qwen3TTSModel.fromModelDirectory (Qwen3TTS.swift:1136-1175): config.json + every *.safetensors ->
        sanitize ("talker." prefix) -> MLX affine de-quantisation as config.json's "quantization" says -> weights on the device, all
        inside the library.  The tokenizer and <dir>/speech_tokenizer (Qwen3TTSSpeechTokenizerDecoder.from_model_directory) stay with
        the caller."""
        self = cls.__new__(cls)
        c = _ffi.Qwen3TalkerConfig()
        _ffi.check(_ffi.lib().b2a_qwen3_talker_config_from_json(str(model_dir).encode() + b"/config.json", max_batch, max_context, C.byref(c)))
        cp = Qwen3CodePredictorConfig(vocab_size=c.cp_vocab_size, hidden_size=c.cp_hidden_size, intermediate_size=c.cp_intermediate_size,
                                      num_hidden_layers=c.cp_num_hidden_layers, num_attention_heads=c.cp_num_attention_heads,
                                      num_key_value_heads=c.cp_num_key_value_heads, head_dim=c.cp_head_dim, rms_norm_eps=c.cp_rms_norm_eps,
                                      rope_theta=c.cp_rope_theta, num_code_groups=c.num_code_groups)
        self.config = Qwen3TalkerConfig(vocab_size=c.vocab_size, hidden_size=c.hidden_size, intermediate_size=c.intermediate_size,
                                        num_hidden_layers=c.num_hidden_layers, num_attention_heads=c.num_attention_heads,
                                        num_key_value_heads=c.num_key_value_heads, head_dim=c.head_dim, rms_norm_eps=c.rms_norm_eps,
                                        rope_theta=c.rope_theta, num_code_groups=c.num_code_groups, text_hidden_size=c.text_hidden_size,
                                        text_vocab_size=c.text_vocab_size, codec_eos_token_id=c.codec_eos_token_id, code_predictor=cp)
        self._h = C.c_void_p()
        _ffi.check(_ffi.lib().b2a_qwen3_talker_create_from_directory(str(model_dir).encode(), device, max_batch, max_context, C.byref(self._h)))
        return self

    @property
    def stream(self) -> int:
        return int(_ffi.lib().b2a_qwen3_talker_stream(self._h) or 0)

    # -- embeddings ------------------------------------------------------------------------------
    def embed_text(self, ids: Sequence[int]) -> np.ndarray:
        """text_projection(text_embedding(ids)) (Qwen3TTS.swift:898) -> [n, hidden]."""
        a = np.ascontiguousarray(ids, dtype=np.int32)
        out = np.empty((len(a), self.config.hidden_size), dtype=np.float32)
        _ffi.check(_ffi.lib().b2a_qwen3_talker_embed_text(self._h, _ffi.ptr(a), len(a), _ffi.ptr(out)))
        return out

    def embed_codec(self, ids: Sequence[int]) -> np.ndarray:
        a = np.ascontiguousarray(ids, dtype=np.int32)
        out = np.empty((len(a), self.config.hidden_size), dtype=np.float32)
        _ffi.check(_ffi.lib().b2a_qwen3_talker_embed_codec(self._h, _ffi.ptr(a), len(a), _ffi.ptr(out)))
        return out

    def prepare_generation_inputs(self, chat_ids: Sequence[int], tts_bos: int, tts_eos: int, tts_pad: int,
                                  language_id: Optional[int] = None, speaker_id: Optional[int] = None,
                                  instruct_ids: Optional[Sequence[int]] = None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """prepareGenerationInputs (Qwen3TTS.swift:883-999) from token ids: chat_ids = tokens of
        "<|im_start|>assistant\\n{text}<|im_end|>\\n<|im_start|>assistant\\n", instruct_ids = tokens of the VoiceDesign instruct turn.
        Returns (input_embeds [L, H], trailing_text_hidden [n, H], tts_pad_embed [H]).  Row selection / concatenation / the two
        adds are host bookkeeping; every embedding row comes from the device."""
        c = self.config
        text = self.embed_text(list(chat_ids))
        tts = self.embed_text([tts_bos, tts_eos, tts_pad])
        bos_e, eos_e, pad_e = tts[0:1], tts[1:2], tts[2:3]
        prefill = ([c.codec_think_id, c.codec_think_bos_id, language_id, c.codec_think_eos_id] if language_id is not None
                   else [c.codec_nothink_id, c.codec_think_bos_id, c.codec_think_eos_id])                          # :938-951
        ids = prefill + ([speaker_id] if speaker_id is not None else []) + [c.codec_pad_id, c.codec_bos_id]          # :957-962
        codec = self.embed_codec(ids)
        pad_count = codec.shape[0] - 2
        combined = np.concatenate([np.repeat(pad_e, pad_count, axis=0), bos_e], axis=0) + codec[:-1]               # :976-979
        pieces = ([self.embed_text(list(instruct_ids))] if instruct_ids else []) + [text[:3], combined]
        first_text = text[3:4] + codec[-1:]                                                                          # :989
        inputs = np.concatenate(pieces + [first_text], axis=0)
        trailing = np.concatenate([text[4:text.shape[0] - 5], eos_e], axis=0)                                        # :993-996
        return inputs.astype(np.float32), trailing.astype(np.float32), pad_e[0].astype(np.float32)

    # -- model -----------------------------------------------------------------------------------
    def __call__(self, input_embeds: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """Talker forward from an empty cache: input_embeds [B, L, H] -> (codec logits of the last position [B, V], hidden [B, H])."""
        x = np.ascontiguousarray(input_embeds, dtype=np.float32)
        if x.ndim == 2:
            x = x[None]
        B, L, H = x.shape
        logits = np.empty((B, self.config.vocab_size), dtype=np.float32)
        hidden = np.empty((B, H), dtype=np.float32)
        _ffi.check(_ffi.lib().b2a_qwen3_talker_forward(self._h, _ffi.ptr(x), B, L, _ffi.ptr(logits), _ffi.ptr(hidden)))
        return logits, hidden

    def generate_codes(self, input_embeds, trailing_text_hidden, tts_pad_embed, parameters: Optional[Qwen3GenerateParameters] = None,
                       on_frame: Optional[Callable[[int, int, np.ndarray], None]] = None):
        """The frame loop (Qwen3TTS.swift:380-495) for B utterances with the same prompt length: input_embeds [B, L, H],
        trailing_text_hidden a list of [n_b, H] arrays (or one [B, n, H] array), tts_pad_embed [H].
        Returns ([codes_b [frames_b, num_code_groups]], AudioGenerationInfo)."""
        p = parameters or Qwen3GenerateParameters()
        x = np.ascontiguousarray(input_embeds, dtype=np.float32)
        if x.ndim == 2:
            x = x[None]
        B, L, H = x.shape
        tr = [np.asarray(t, dtype=np.float32).reshape(-1, H) for t in (trailing_text_hidden if not isinstance(trailing_text_hidden, np.ndarray)
                                                                      or trailing_text_hidden.ndim == 3 else [trailing_text_hidden])]
        if len(tr) != B:
            raise _ffi.AudioGenerationError(_ffi.ERR_INVALID_INPUT, "one trailing-text block per utterance")
        nmax = max((t.shape[0] for t in tr), default=0)
        trail = np.zeros((B, max(nmax, 1), H), dtype=np.float32)
        nt = np.zeros(B, dtype=np.int32)
        for b, t in enumerate(tr):
            trail[b, :t.shape[0]] = t
            nt[b] = t.shape[0]
        pad = np.ascontiguousarray(tts_pad_embed, dtype=np.float32).reshape(H)
        G = self.config.num_code_groups
        codes = np.zeros((B, p.max_tokens, G), dtype=np.int32)
        nfr = np.zeros(B, dtype=np.int32)
        info = _ffi.GenInfo()
        gp = p._c()
        _ffi.check(_ffi.lib().b2a_qwen3_talker_set_bench_flags(self._h, int(p.mask_eos)))
        cb = (FRAME_CB(lambda user, b, f, c: on_frame(b, f, np.ctypeslib.as_array(c, shape=(G,)).copy())) if on_frame else FRAME_CB())
        _ffi.check(_ffi.lib().b2a_qwen3_talker_generate(self._h, _ffi.ptr(x), B, L, _ffi.ptr(trail), _ffi.ptr(nt), nmax, _ffi.ptr(pad), C.byref(gp),
                                                        _ffi.ptr(codes), _ffi.ptr(nfr), C.byref(info), cb, None))
        gi = AudioGenerationInfo(info.prompt_token_count, info.generation_token_count, info.prefill_time, info.generate_time,
                                 info.tokens_per_second, info.peak_memory_gb, info.codec_time)
        return [codes[b, :nfr[b]].copy() for b in range(B)], gi

    def cancel(self) -> None:
        _ffi.check(_ffi.lib().b2a_qwen3_talker_cancel(self._h))

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                _ffi.lib().b2a_qwen3_talker_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:   # interpreter shutdown
            pass


class Qwen3TTSModel:
    """SpeechGenerationModel face of Qwen3-TTS (Qwen3TTS.swift:306-569): talker + code predictor -> codes -> speech-tokenizer
    decoder.  `speech_tokenizer` is a qwen3_tts_codec.Qwen3TTSSpeechTokenizer (the decode side; borrowed)."""
    sample_rate = 24000

    def __init__(self, talker: Qwen3TTSTalker, speech_tokenizer=None):
        self.talker, self.speech_tokenizer = talker, speech_tokenizer

    def generate(self, input_embeds, trailing_text_hidden, tts_pad_embed, parameters: Optional[Qwen3GenerateParameters] = None) -> np.ndarray:
        """generate (:412-510) after prepare_generation_inputs: one utterance -> 1-D waveform (chunked decode, :1059-1068)."""
        if self.speech_tokenizer is None:
            raise _ffi.AudioGenerationError(_ffi.ERR_MODEL_NOT_INITIALIZED, "speech tokenizer not loaded")
        codes, _ = self.talker.generate_codes(np.asarray(input_embeds)[None], [trailing_text_hidden], tts_pad_embed, parameters)
        if codes[0].shape[0] == 0:
            raise _ffi.AudioGenerationError(_ffi.ERR_GENERATION_FAILED, "No audio codes generated")
        wav, lengths = self.speech_tokenizer.decode(codes[0][None])
        return wav[0, :int(lengths[0])]

    def generate_stream(self, input_embeds, trailing_text_hidden, tts_pad_embed, parameters: Optional[Qwen3GenerateParameters] = None,
                        streaming_interval: float = 2.0) -> Iterator:
        """generateStream (:512-569): audio chunks DURING generation -- every int(streaming_interval * 12.5) frames the new codes go
        through the speech tokenizer's streaming step (decodeChunk -> streamingDecode, Qwen3TTS.swift:214-231) and are yielded as
        ('audio', samples); ('token', c0) per frame, ('info', ...) at the end."""
        if self.speech_tokenizer is None:
            raise _ffi.AudioGenerationError(_ffi.ERR_MODEL_NOT_INITIALIZED, "speech tokenizer not loaded")
        chunk = max(1, int(streaming_interval * 12.5))
        dec = self.speech_tokenizer.decoder
        dec.reset_streaming_state()
        events: List = []
        pending: List[np.ndarray] = []

        def flush():
            if pending:
                c = np.stack(pending)[None].transpose(0, 2, 1)                # [1, G, n]
                events.append(("audio", dec.streaming_step(np.ascontiguousarray(c))[0, 0]))
                pending.clear()

        def on_frame(b, f, c):
            events.append(("token", int(c[0])))
            pending.append(c)
            if len(pending) >= chunk:
                flush()

        _, info = self.talker.generate_codes(np.asarray(input_embeds)[None], [trailing_text_hidden], tts_pad_embed, parameters, on_frame=on_frame)
        flush()
        yield from events
        yield ("info", info)
