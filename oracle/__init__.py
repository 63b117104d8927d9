"""CPU oracle: a restatement of the reference's algorithm for the hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import anything from this package.  The product path
(``mlx-audio-swift_b200``) never imports it and fails loudly when its CUDA
library is missing.

PARITY UNPINNED.  The reference (Blaizzy/mlx-audio-swift @ 4266f988) is pure
Swift on top of the un-vendored ``mlx-swift 0.31.4`` / ``mlx-swift-lm 3.31.4``
packages (``Package.resolved``); neither can be compiled or imported in this
container (no Swift toolchain, macOS/iOS-only package, no ``mlx`` Python
module), and the reference's own tests hold no numeric golden vectors for this
path -- only two "returns nil" robustness cases for
``IncrementalMelSpectrogram`` (``Tests/IncrementalMelSpectrogramTests.swift:7-17``),
shape checks for Whisper features (``Tests/MLXAudioSTTTests.swift:4416-4422``)
and closed-form window values.  Those are all reproduced in
``tests/test_oracle_*.py``.  Beyond that the oracle is cross-checked against independent implementations available offline
(``tests/test_oracle_*.py``): ``torch.stft`` and ``torchaudio`` (mel filterbanks, the whole offline log-mel), ``transformers``
Whisper (feature extractor, encoder / decoder, sinusoids), Llama with llama3 rope scaling, the logits processors (repetition
penalty, top-k / top-p / min-p), EncodecModel (decoder, RVQ, linear overlap-add), DAC (the SNAC decoder in its dense mode, Snake,
the cosine code search), ConvNext (Vocos block), ``torch.istft`` (Vocos head), Qwen3 / Qwen3-VL / Qwen3-Omni Code2Wav (Qwen3-TTS).

Modules: ``dsp`` (mel), ``snac``, ``llama`` (Orpheus), ``whisper``, ``vocos``, ``encodec`` -- the rows of SURVEY.md section 8a --
and ``qwen3_tts`` + ``qwen3_tts_codec`` (row N1 of 8f, talker / code predictor and speech-tokenizer decoder: oracle
only, no CUDA path yet).

Every function cites the reference ``file:line`` it follows (paths relative
to the reference checkout root).
"""
