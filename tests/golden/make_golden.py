"""Generates the committed golden fixtures from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference itself cannot run here, so
these pin the ORACLE; the oracle in turn is cross-checked against independent implementations in
tests/test_oracle_*.py.  Style follows the reference's own Python-parity goldens: first-N samples +
mean / abs-mean / min / max (Tests/MLXAudioCodecsTests.swift:207-241), plus small full arrays."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import dsp, llama, snac  # noqa: E402
from oracle import encodec as oe  # noqa: E402
from oracle import qwen3_tts_codec as oq  # noqa: E402
from oracle import qwen3_tts as ot  # noqa: E402
from oracle import vocos as ov  # noqa: E402
from oracle import whisper as ow  # noqa: E402

OUT = Path(__file__).resolve().parent


def stats(x):
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    return np.array([x.mean(), np.abs(x).mean(), x.min(), x.max()])


def mel():
    x = dsp.synth_audio(160000, 0)
    m = dsp.IncrementalMelSpectrogram(16000, 400, 160, 80)
    a, b = m.process(x), m.flush()
    full = np.concatenate([a, b])
    # irregular chunking (the reference's streaming==offline test style)
    m2 = dsp.IncrementalMelSpectrogram(16000, 400, 160, 80)
    cuts = [0, 1, 3, 150, 700, 5000, 5160, 40000, 160000]
    chunks = []
    for i in range(len(cuts) - 1):
        o = m2.process(x[cuts[i]:cuts[i + 1]])
        if o is not None:
            chunks.append(o)
    chunks.append(m2.flush())
    irr = np.concatenate(chunks)
    w = dsp.whisper_encoder_features(x, 80)[0]
    off = dsp.compute_mel_spectrogram(x, 16000, 400, 160, 80)
    np.savez_compressed(OUT / "mel.npz", inc_first=full[:4].astype(np.float32), inc_last=full[-3:].astype(np.float32),
                        inc_stats=stats(full), inc_shape=np.array(full.shape), irr_stats=stats(irr),
                        irr_shape=np.array(irr.shape), irr_rows=irr[[0, 5, 30, 500, 1000]].astype(np.float32),
                        whisper_stats=stats(w), whisper_rows=w[[0, 1, 999, 1000, 2999]].astype(np.float32),
                        core_stats=stats(off), core_rows=off[[0, 1, 500, 1000]].astype(np.float32))


def snac_small():
    cfg = snac.SNACConfig()
    W = snac.init_weights(cfg, 1234)
    codes = snac.synth_codes(cfg, 2, 16, seed=2)
    rng = np.random.default_rng(7)
    noise = [rng.standard_normal(s).astype(np.float32) for s in snac.noise_shapes(cfg, 2, 16)]
    y = snac.decode(cfg, W, codes, noise)
    y0 = snac.decode(cfg, W, codes, None)
    z = (rng.standard_normal((2, cfg.latent, 16)) * 0.5).astype(np.float32)
    zq, qc = snac.quantize(cfg, W, z)
    np.savez_compressed(OUT / "snac.npz", wave=y.astype(np.float32), wave_nonoise=y0.astype(np.float32),
                        wave_stats=stats(y), q_codes0=qc[0], q_codes1=qc[1], q_codes2=qc[2], zq_stats=stats(zq))


def llama_tiny():
    cfg = llama.LlamaConfig(hidden_size=256, num_hidden_layers=2, intermediate_size=512, num_attention_heads=2,
                            num_key_value_heads=1, head_dim=128, vocab_size=2048)
    W = llama.init_weights(cfg, 1234, std=0.08)
    ids = np.random.default_rng(3).integers(0, 2048, size=(2, 12))
    mo = llama.LlamaOracle(cfg, W, round_acts=False)
    lg = mo.forward(torch.as_tensor(ids)).numpy()
    gen = llama.generate_tokens(llama.LlamaOracle(cfg, W, False), ids, max_tokens=24, temperature=0.0,
                                rep_penalty=1.3, rep_context=20)
    np.savez_compressed(OUT / "llama_tiny.npz", ids=ids, logits_last=lg[:, -1].astype(np.float32),
                        logits_stats=stats(lg), greedy=np.asarray(gen, dtype=np.int32),
                        freqs=llama.llama3_rope_freqs(llama.LlamaConfig()))


def whisper_tiny():
    cfg = ow.WhisperConfig.tiny_test()
    W = ow.init_weights(cfg, 1234)
    x = dsp.synth_audio(64000, 3)
    o = ow.WhisperOracle(cfg, W)
    feats = torch.from_numpy(dsp.whisper_encoder_features(x)).float()
    enc = o.encode(feats).numpy()
    toks, logits = ow.transcribe_tokens(ow.WhisperOracle(cfg, W), x, ow.build_prompt_tokens(), max_tokens=12, mask_eot=True,
                                        return_logits=True)
    np.savez_compressed(OUT / "whisper_tiny.npz", greedy=np.asarray(toks, dtype=np.int32), enc_stats=stats(enc),
                        enc_rows=enc[0, [0, 1, 700, 1499]].astype(np.float32), first_logits_stats=stats(np.clip(logits[0], -50, 50)))


def codecs_small():
    """Vocos (reference test geometry, 2 layers to keep it small) and Encodec-24 kHz decode: first-N + stats."""
    vc = ov.VocosConfig(num_layers=2)
    VW = ov.init_weights(vc, 7)
    f = np.random.default_rng(1).standard_normal((2, 37, vc.input_channels)).astype(np.float32)
    y = ov.decode(vc, VW, f)
    ec = oe.EncodecConfig()
    EW = oe.init_weights(ec, 7, n_codebooks=8)
    codes = np.random.default_rng(1).integers(0, 1024, size=(1, 3, 8, 41))
    z = oe.decode(ec, EW, codes)
    np.savez_compressed(OUT / "codecs.npz", vocos_first=y[:, :64].astype(np.float32), vocos_stats=stats(y), vocos_shape=np.array(y.shape),
                        encodec_first=z[:, :64, 0].astype(np.float32), encodec_last=z[:, -64:, 0].astype(np.float32),
                        encodec_stats=stats(z), encodec_shape=np.array(z.shape))


def qwen3_codec_config():
    """The mid-size geometry the (gated) GPU tests of row N1 use."""
    return oq.mid_config()


def qwen3_codec():
    """Qwen3-TTS speech-tokenizer decoder (row N1): one-shot decode and a three-chunk streaming decode (bias-twice behaviour included)."""
    cfg = qwen3_codec_config()
    W = oq.init_weights(cfg, 5)
    codes = np.random.default_rng(1).integers(0, cfg.codebook_size, (2, cfg.num_quantizers, 20))
    d = oq.SpeechTokenizerDecoder(cfg, W)
    full = d(codes)[:, 0].numpy()
    d.reset_streaming_state()
    st = np.concatenate([d.streaming_step(codes[:, :, a:b])[:, 0].numpy() for a, b in ((0, 7), (7, 8), (8, 20))], axis=-1)
    up = cfg.total_upsample
    np.savez_compressed(OUT / "qwen3_codec.npz", full_first=full[:, :64].astype(np.float32), full_last=full[:, -64:].astype(np.float32), full_stats=stats(full),
                        stream_boundary=st[:, 7 * up - 32: 8 * up + 32].astype(np.float32), stream_stats=stats(st), shape=np.array(full.shape))


def qwen3_talker_config():
    """The geometry of the reference's own Qwen3-TTS tests (Tests/MLXAudioTTSTests.swift:615-687) with 4 code groups."""
    cp = ot.CodePredictorConfig(vocab_size=2048, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                                head_dim=8, num_code_groups=4)
    return ot.TalkerConfig(vocab_size=3072, hidden_size=32, intermediate_size=64, num_hidden_layers=3, num_attention_heads=4, num_key_value_heads=2,
                           head_dim=8, num_code_groups=4, text_hidden_size=48, text_vocab_size=200, codec_eos_token_id=2150, mrope_section=(2, 1, 1),
                           code_predictor=cp)


QWEN3_CHAT_IDS = [151, 12, 13, 40, 41, 42, 43, 44, 45, 152, 14, 151, 12, 13]


def qwen3_talker_weights(cfg):
    """bf16-VALUED weights (what a checkpoint holds and what a bf16-weight engine can represent exactly), norm gains included."""
    return {k: v.to(torch.bfloat16).to(torch.float32) for k, v in ot.init_weights(cfg, 3).items()}


def qwen3_talker():
    """Qwen3-TTS talker + code predictor (row N1): prompt embeddings, first-step logits and 5 greedy frames (T = 0: no RNG involved)."""
    cfg = qwen3_talker_config()
    W = qwen3_talker_weights(cfg)
    inp, trail, pad = ot.prepare_generation_inputs(cfg, W, QWEN3_CHAT_IDS, tts_bos=160, tts_eos=161, tts_pad=162, language_id=2160)
    logits, hidden = ot.Talker(cfg, W)(inp, None)
    codes = ot.generate_codes(cfg, W, inp, trail, pad, max_tokens=5, temperature=0.0, repetition_penalty=1.05, stop_on_eos=False)
    np.savez_compressed(OUT / "qwen3_talker.npz", input_embeds_stats=stats(inp.numpy()), input_shape=np.array(inp.shape),
                        first_logits_stats=stats(logits[0, -1].numpy()), first_logits_top=np.argsort(-logits[0, -1].numpy())[:8].astype(np.int32),
                        codes=codes.numpy().astype(np.int32))


def qwen3_talker_hd128():
    """The same, at the smallest geometry the CUDA engine runs (head_dim 128): 5 greedy frames of tests/test_gpu_qwen3_talker.py's
    small model (bf16-valued weights, T = 0)."""
    cp = ot.CodePredictorConfig(vocab_size=2048, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                                num_key_value_heads=1, head_dim=128, num_code_groups=4)
    cfg = ot.TalkerConfig(vocab_size=3072, hidden_size=256, intermediate_size=512, num_hidden_layers=3, num_attention_heads=2,
                          num_key_value_heads=1, head_dim=128, num_code_groups=4, text_hidden_size=128, text_vocab_size=200,
                          codec_eos_token_id=2150, code_predictor=cp)
    W = {k: v.to(torch.bfloat16).to(torch.float64) for k, v in ot.init_weights(cfg, 3, std=0.05).items()}
    chat = [151, 12, 13, 40, 41, 42, 43, 44, 45, 46, 47, 152, 14, 151, 12, 13]
    inp, trail, pad = ot.prepare_generation_inputs(cfg, W, chat, tts_bos=160, tts_eos=161, tts_pad=162, language_id=2160)
    logits, _ = ot.Talker(cfg, W)(inp, None)
    codes = ot.generate_codes(cfg, W, inp, trail, pad, max_tokens=5, temperature=0.0, repetition_penalty=1.05, stop_on_eos=False)
    np.savez_compressed(OUT / "qwen3_talker_hd128.npz", first_logits_stats=stats(logits[0, -1].numpy()),
                        first_logits_top=np.argsort(-logits[0, -1].numpy())[:8].astype(np.int32), codes=codes.numpy().astype(np.int32))


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None, help="regenerate a single fixture (mel | snac | llama | whisper | codecs | qwen3_codec | qwen3_talker | qwen3_talker_hd128)")
    only = ap.parse_args().only
    for name, fn in (("mel", mel), ("snac", snac_small), ("llama", llama_tiny), ("whisper", whisper_tiny), ("codecs", codecs_small), ("qwen3_codec", qwen3_codec), ("qwen3_talker", qwen3_talker), ("qwen3_talker_hd128", qwen3_talker_hd128)):
        if only is None or only == name:
            fn()
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size)
