"""Oracle Whisper restatement vs transformers' WhisperForConditionalGeneration (random init, same weights), the
reference's prompt / suppression semantics, and committed goldens.  CPU only."""
import numpy as np
import torch

from conftest import GOLDEN, rel_err
from oracle import dsp
from oracle import whisper as ow


def _hf(cfg, W):
    from transformers import WhisperConfig as HC, WhisperForConditionalGeneration
    hc = HC(vocab_size=cfg.vocab_size, num_mel_bins=cfg.num_mel_bins, d_model=cfg.d_model, encoder_layers=cfg.encoder_layers,
            encoder_attention_heads=cfg.encoder_attention_heads, encoder_ffn_dim=cfg.encoder_ffn_dim,
            decoder_layers=cfg.decoder_layers, decoder_attention_heads=cfg.decoder_attention_heads,
            decoder_ffn_dim=cfg.decoder_ffn_dim, max_source_positions=1500, max_target_positions=448,
            attn_implementation="eager")
    m = WhisperForConditionalGeneration(hc).float().eval()
    sd = {k: v.float() for k, v in W.items()}
    sd["proj_out.weight"] = sd["model.decoder.embed_tokens.weight"]
    m.load_state_dict(sd, strict=True)
    return m


def test_encoder_and_decoder_match_transformers():
    cfg = ow.WhisperConfig.tiny_test()
    W = ow.init_weights(cfg, 1)
    m = _hf(cfg, W)
    feats = torch.from_numpy(dsp.whisper_encoder_features(dsp.synth_audio(48000, 0))).float()
    o = ow.WhisperOracle(cfg, W)
    enc = o.encode(feats)
    with torch.no_grad():
        ref_enc = m.model.encoder(feats.transpose(1, 2)).last_hidden_state
    assert enc.shape == (1, 1500, cfg.d_model) and rel_err(enc.numpy(), ref_enc.numpy()) < 1e-5
    ids = torch.as_tensor([ow.build_prompt_tokens() + [100, 200, 300]])
    lg = o.logits(o.decode(ids, 0, enc))
    with torch.no_grad():
        ref = m(input_features=feats.transpose(1, 2), decoder_input_ids=ids).logits
    assert rel_err(lg.numpy(), ref.numpy()) < 1e-5
    # incremental decode with the self-attention KV cache == full pass (WhisperLayers.swift:49-55)
    o.reset()
    a = o.logits(o.decode(ids[:, :4], 0, enc))
    b = torch.cat([o.logits(o.decode(ids[:, 4 + i:5 + i], 4 + i, enc)) for i in range(3)], dim=1)
    assert rel_err(torch.cat([a, b], dim=1).numpy(), lg.numpy()) < 1e-5


def test_prompt_tokens_and_suppression():
    assert ow.build_prompt_tokens() == [50258, 50259, 50359, 50363]                    # <|sot|><|en|><|transcribe|><|notimestamps|>
    assert ow.build_prompt_tokens(None, "translate") == [50258, 50358, 50363]
    assert ow.build_prompt_tokens(multilingual=False) == [50258, 50363]
    cfg = ow.WhisperConfig.tiny_test()
    W = ow.init_weights(cfg, 2)
    x = dsp.synth_audio(32000, 1)
    toks, logits = ow.transcribe_tokens(ow.WhisperOracle(cfg, W), x, ow.build_prompt_tokens(), max_tokens=6, return_logits=True)
    assert all(t < ow.TIMESTAMP_BEGIN for t in toks)                                     # timestamps always suppressed
    assert logits[0][ow.EOT] < -1e8 and (len(logits) < 2 or logits[1][ow.EOT] > -1e8)    # EOT suppressed at step 0 only
    forced = ow.transcribe_tokens(ow.WhisperOracle(cfg, W), x, ow.build_prompt_tokens(), max_tokens=500, mask_eot=True)
    assert len(forced) == 448 - 4 - 1                                                    # min(maxTokens, maxTarget - prompt - 1)


def test_sinusoids_match_reference_formula():
    s = ow.sinusoids(1500, 512).numpy()
    assert s.shape == (1500, 512) and abs(s[0, 0]) < 1e-7 and abs(s[0, 256] - 1) < 1e-7
    assert abs(s[10, 255] - np.sin(10 * np.exp(-np.log(10000.0)))) < 1e-6


def test_goldens():
    g = np.load(GOLDEN / "whisper_tiny.npz")
    cfg = ow.WhisperConfig.tiny_test()
    W = ow.init_weights(cfg, 1234)
    x = dsp.synth_audio(64000, 3)
    toks = ow.transcribe_tokens(ow.WhisperOracle(cfg, W), x, ow.build_prompt_tokens(), max_tokens=12, mask_eot=True)
    assert toks == g["greedy"].tolist()


def test_sinusoids_match_transformers():
    """whisperSinusoids (WhisperModel.swift:384-397) against transformers' own `sinusoids` (what HF Whisper checkpoints are initialised with)."""
    from transformers.models.whisper.modeling_whisper import sinusoids as hf_sinusoids
    for length, channels in ((1500, 512), (1500, 384), (32, 64)):
        # the reference evaluates sin / cos in Double and then narrows (:387-395); transformers works in float32, so the angle of the
        # late positions carries ~1e-4 of rounding there -- same table up to that
        assert np.abs(ow.sinusoids(length, channels).numpy() - hf_sinusoids(length, channels).numpy()).max() < 2e-4
