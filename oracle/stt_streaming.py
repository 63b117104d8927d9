"""TEST INFRASTRUCTURE (CPU oracle): the reference's streaming STT session for a generic `STTGenerationModel`, token level.

Restates `Sources/MLXAudioSTT/Streaming/StreamingInferenceSession.swift:589-950` (the core that drives a model through
`streamingDecodeTokenIds(audio:config:confirmedTokenIds:)`) with the clock passed in and the decode function injected:

* `feed(samples, now)` (:589-638): samples are appended to the pending buffer.  If the buffer holds a whole window (8 s in the
  reference), the window is frozen -- the buffer keeps its last `overlap` seconds -- and decoded once, with no prefix
  (`finalWindow`, :727-748): its tokens join the completed list and the confirmed / provisional state is cleared.  Otherwise, if at
  least half a second is pending and `max(0.2, decode_interval)` seconds have passed since the last pass, the WHOLE pending buffer is
  decoded with the confirmed tokens as a forced prefix (`partial`).
* `promote` (:750-829): the tokens after the confirmed prefix are the new provisional list; a position keeps its first-seen time and
  gains one agreement while it repeats the previous pass's token at that position (longest common prefix); the longest prefix whose
  every token is older than the delay preset AND agreed on by `min_agreement_passes` passes moves to the confirmed list.
* `stop(now)` (:831-947): what is still pending is decoded as a final window; left-over provisional tokens become confirmed.

Text (de-duplication of the window overlap, tokenizer) stays with the host; this module and the CUDA library's session work on ids.
`decode(audio: np.ndarray, prefix: list[int]) -> list[int]` returns the continuation AFTER the prefix."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np


@dataclass
class StreamingConfig:          # StreamingTypes.swift:36-92 (the fields this core reads)
    decode_interval_s: float = 1.0
    window_s: float = 8.0
    window_overlap_s: float = 1.0
    delay_ms: int = 480          # DelayPreset.agent
    min_agreement_passes: int = 2
    sample_rate: int = 16000


@dataclass
class Update:
    kind: str                    # "none" | "partial" | "final_window" | "ended"
    promoted: int = 0
    completed: List[List[int]] = field(default_factory=list)
    confirmed: List[int] = field(default_factory=list)
    provisional: List[int] = field(default_factory=list)
    finalized_windows: int = 0
    total_audio_s: float = 0.0


class StreamingSession:
    def __init__(self, decode: Callable[[np.ndarray, List[int]], List[int]], config: Optional[StreamingConfig] = None):
        self.decode, self.cfg = decode, config or StreamingConfig()
        c = self.cfg
        self.window = int(c.sample_rate * c.window_s)
        self.overlap = max(0, min(int(round(c.window_overlap_s * c.sample_rate)), max(0, self.window - 1)))     # :578-584
        self.pending = np.zeros(0, np.float32)
        self.total = 0
        self.last_decode: Optional[float] = None
        self.finalized = 0
        self.completed: List[List[int]] = []
        self.confirmed: List[int] = []
        self.provisional: List[int] = []
        self.first_seen: List[float] = []
        self.agreement: List[int] = []
        self.active = True

    def _update(self, kind: str, promoted: int = 0) -> Update:
        return Update(kind, promoted, [list(w) for w in self.completed], list(self.confirmed), list(self.provisional), self.finalized,
                      self.total / self.cfg.sample_rate)

    def _finalize(self, audio: np.ndarray) -> None:                      # finalizeWindow :727-748
        self.completed.append(list(self.decode(audio, [])))
        self.confirmed, self.provisional, self.first_seen, self.agreement = [], [], [], []

    def _promote(self, new: List[int], now: float) -> int:               # promoteTokens :750-829
        delay = self.cfg.delay_ms / 1000.0
        match = 0
        for a, b in zip(self.provisional, new):
            if a != b:
                break
            match += 1
        seen, agree = [], []
        for i in range(len(new)):
            if i < match:
                seen.append(self.first_seen[i] if i < len(self.first_seen) else now)
                agree.append(max(1, (self.agreement[i] if i < len(self.agreement) else 1) + 1))
            else:
                seen.append(now)
                agree.append(1)
        promote = 0
        need = max(1, self.cfg.min_agreement_passes)
        for i in range(len(new)):
            if now - seen[i] >= delay and agree[i] >= need:
                promote = i + 1
            else:
                break
        self.confirmed += new[:promote]
        self.provisional, self.first_seen, self.agreement = new[promote:], seen[promote:], agree[promote:]
        return promote

    def feed(self, samples, now: float) -> Update:                       # feedAudio :589-638
        if not self.active:
            return self._update("none")
        x = np.asarray(samples, np.float32).reshape(-1)
        self.pending = np.concatenate([self.pending, x])
        self.total += len(x)
        if len(self.pending) >= self.window:
            win = self.pending[:self.window]
            self.pending = self.pending[max(0, self.window - self.overlap):]
            self.finalized += 1
            self.last_decode = now
            self._finalize(win)
            return self._update("final_window")
        if len(self.pending) < self.cfg.sample_rate // 2:
            return self._update("none")
        if self.last_decode is not None and now - self.last_decode < max(0.2, self.cfg.decode_interval_s):
            return self._update("none")
        self.last_decode = now
        new = list(self.decode(self.pending, list(self.confirmed)))
        return self._update("partial", self._promote(new, now))

    def stop(self, now: float) -> Update:                                # stop / finishStop :831-947
        if not self.active:
            return self._update("none")
        self.active = False
        if len(self.pending):
            self.finalized += 1
            self._finalize(self.pending)
        self.confirmed += self.provisional
        self.provisional, self.first_seen, self.agreement = [], [], []
        self.pending = np.zeros(0, np.float32)
        return self._update("ended")
