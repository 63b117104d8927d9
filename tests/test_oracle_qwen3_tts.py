"""Oracle for the Qwen3-TTS talker / code predictor pinned against independent implementations available
offline: `transformers` Qwen3Model (the talker backbone: per-head q/k RMSNorm, rotate-half RoPE, GQA, SwiGLU) and Qwen3-VL's
apply_interleaved_mrope; plus the closed-form semantics of sampleToken and the frame loop's cache discipline.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import qwen3_tts as oq

SMALL = dict(vocab_size=1100, hidden_size=64, intermediate_size=96, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
             head_dim=16, text_hidden_size=48, text_vocab_size=200, codec_eos_token_id=1090, mrope_section=(4, 2, 2))


def small_cfg(**kw):
    cp = oq.CodePredictorConfig(vocab_size=40, hidden_size=kw.pop("cp_hidden", 64), intermediate_size=80, num_hidden_layers=2,
                                num_attention_heads=4, num_key_value_heads=2, head_dim=16, num_code_groups=kw.pop("groups", 5))
    d = dict(SMALL); d.update(kw)
    return oq.TalkerConfig(num_code_groups=cp.num_code_groups, code_predictor=cp, **d)


def test_talker_backbone_matches_transformers_qwen3():
    from transformers import Qwen3Config, Qwen3Model
    cfg = small_cfg()
    W = oq.init_weights(cfg, 3, std=0.2)
    hc = Qwen3Config(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                     num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                     num_key_value_heads=cfg.num_key_value_heads, head_dim=cfg.head_dim, rms_norm_eps=cfg.rms_norm_eps,
                     rope_theta=cfg.rope_theta, attention_bias=False, max_position_embeddings=512, tie_word_embeddings=False,
                     use_sliding_window=False, attn_implementation="eager")
    m = Qwen3Model(hc).double().eval()
    sd = m.state_dict()
    for k in sd:
        src = "model." + k if not k.startswith("embed_tokens") else "model.codec_embedding.weight"
        sd[k] = W[src].to(torch.float64)
    m.load_state_dict(sd)
    x = torch.randn(2, 7, cfg.hidden_size, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        ref = m(inputs_embeds=x, use_cache=False).last_hidden_state
    mine = oq.Talker(cfg, W).model(x)
    assert (mine - ref).abs().max() < 1e-5 * ref.abs().max()          # HF evaluates the rotary angles in float32 (2e-7 here)
    # incremental decoding through the KV cache == the full pass
    t = oq.Talker(cfg, W)
    cache = t.make_cache()
    a = t.model(x[:, :4], cache)
    outs = [a] + [t.model(x[:, i:i + 1], cache) for i in range(4, 7)]
    assert (torch.cat(outs, dim=1) - mine).abs().max() < 1e-10


def test_interleaved_mrope_matches_qwen3_vl():
    from transformers.models.qwen3_vl.modeling_qwen3_vl import Qwen3VLTextRotaryEmbedding
    freqs = torch.randn(3, 2, 5, 64, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    for section in ((24, 20, 20), (16, 24, 24), (40, 12, 12)):
        ref = Qwen3VLTextRotaryEmbedding.apply_interleaved_mrope(None, freqs.clone(), list(section))
        assert torch.equal(oq.apply_interleaved_mrope(freqs, section), ref)
    # identical position rows: MRoPE degenerates to plain RoPE (what the TTS text-only positions are, Qwen3TTSTalker.swift:281-285)
    pos = torch.arange(3, 9).view(1, 6)
    c1, s1 = oq.mrope_cos_sin(pos, 128, 1e6, (24, 20, 20))
    c2, s2 = oq.rope_cos_sin(pos, 128, 1e6)
    assert torch.equal(c1, c2) and torch.equal(s1, s2)


def test_sample_token_filters():
    V = 12
    lg = torch.tensor([[2.0, 1.0, 0.5, 0.0, -0.5, -1.0, -1.5, -2.0, 3.0, -3.0, 0.25, 0.75]], dtype=torch.float64)
    # suppress -> -inf; repetition penalty over UNIQUE generated tokens (positive / by p, negative * p)
    f = oq.filter_logits(lg, temperature=0.0, repetition_penalty=2.0, generated_tokens=[0, 0, 4, 99], suppress_tokens=[8])
    assert f[0, 8] == float("-inf") and f[0, 0] == 1.0 and f[0, 4] == -1.0 and f[0, 1] == 1.0
    assert int(oq.sample_token(lg, temperature=0.0, suppress_tokens=[8])[0, 0]) == 0
    # top-k keeps exactly the k largest
    f = oq.filter_logits(lg, temperature=1.0, top_k=3)
    assert set(torch.isfinite(f[0]).nonzero().flatten().tolist()) == {8, 0, 1}
    # top-p: ascending cumulative mass > 1 - p survives (the reference's comparison, Qwen3TTS.swift:1080-1084)
    p = torch.softmax(lg, -1)[0]
    order = torch.argsort(lg[0])
    cum = torch.cumsum(p[order], 0)
    keep = set(order[cum > 1 - 0.6].tolist())
    f = oq.filter_logits(lg, temperature=1.0, top_k=0, top_p=0.6)
    assert set(torch.isfinite(f[0]).nonzero().flatten().tolist()) == keep
    # min-p: logits below max + log(min_p) are dropped
    f = oq.filter_logits(lg, temperature=1.0, top_k=0, min_p=0.2)
    assert set(torch.isfinite(f[0]).nonzero().flatten().tolist()) == set((lg[0] >= 3.0 + np.log(0.2)).nonzero().flatten().tolist())
    # the EOS logit is put back after the filters (so top-k cannot make stopping impossible)
    f = oq.filter_logits(lg, temperature=1.0, top_k=2, eos_token_id=9)
    assert f[0, 9] == -3.0 and torch.isfinite(f[0]).sum() == 3
    g = torch.Generator().manual_seed(0)
    draws = {int(oq.sample_token(lg, g, temperature=1.0, top_k=2)[0, 0]) for _ in range(50)}
    assert draws <= {8, 0} and len(draws) == 2


@pytest.mark.parametrize("cp_hidden", [64, 32])
def test_frame_loop_shapes_cache_discipline_and_determinism(cp_hidden):
    cfg = small_cfg(cp_hidden=cp_hidden)
    W = oq.init_weights(cfg, 5, std=0.3)
    assert ("code_predictor.small_to_mtp_projection.weight" in W) == (cp_hidden != cfg.hidden_size)
    t = oq.Talker(cfg, W)
    prompt = t.embed_text(torch.tensor([[3, 7, 11, 19]])) + t.embed_codec(torch.tensor([[1, 2, 3, 4]]))
    trailing = t.embed_text(torch.tensor([[5, 6]]))
    pad = t.embed_text(torch.tensor([[0]]))
    codes = oq.generate_codes(cfg, W, prompt, trailing, pad, max_tokens=6, temperature=0.0, stop_on_eos=False)
    assert codes.shape == (6, cfg.num_code_groups)
    first = codes[:, 0]                                                     # the last 1024 ids (special tokens) are suppressed, EOS aside
    assert bool(((first < cfg.vocab_size - 1024) | (first == cfg.codec_eos_token_id)).all())
    assert int(codes[:, 1:].max()) < cfg.code_predictor.vocab_size
    again = oq.generate_codes(cfg, W, prompt, trailing, pad, max_tokens=6, temperature=0.0, stop_on_eos=False)
    assert torch.equal(codes, again)
    # frame f's 15 predictor codes depend only on (talker hidden, code 0) of that frame: recompute frame 0 with a fresh predictor
    pred = oq.CodePredictor(cfg, W)
    logits, hidden = t(prompt, t.make_cache())
    suppress = [i for i in range(cfg.vocab_size - 1024, cfg.vocab_size) if i != cfg.codec_eos_token_id]
    c0 = oq.sample_token(logits[:, -1], temperature=0.0, suppress_tokens=suppress)
    cache = pred.make_cache()
    got = [c0]
    for ci in range(cfg.num_code_groups - 1):
        inp = torch.cat([hidden[:, -1:], t.embed_codec(c0)], dim=1) if ci == 0 else pred.embed(ci - 1, got[-1])
        got.append(pred(inp, cache, ci)[:, -1].argmax(-1, keepdim=True))
    assert torch.equal(torch.cat(got, dim=1)[0], codes[0])
    # sanitize strips the "talker." prefix and drops everything else (Qwen3TTSTalker.swift:356-365)
    s = oq.sanitize({"talker.codec_head.weight": W["codec_head.weight"], "speaker_encoder.x": W["codec_head.weight"]})
    assert list(s) == ["codec_head.weight"]


def test_default_geometry_matches_the_reference_defaults():
    c = oq.TalkerConfig()
    assert (c.hidden_size, c.num_hidden_layers, c.num_attention_heads, c.num_key_value_heads, c.head_dim, c.intermediate_size,
            c.vocab_size, c.text_vocab_size, c.text_hidden_size, c.codec_eos_token_id) == (1024, 28, 16, 8, 128, 3072, 3072, 151936, 2048, 2150)
    p = c.code_predictor
    assert (p.num_hidden_layers, p.hidden_size, p.vocab_size, p.num_code_groups) == (5, 1024, 2048, 16)
    assert tuple(c.mrope_section) == (24, 20, 20) and sum(c.mrope_section) == c.head_dim // 2


def test_prepare_generation_inputs_layout():
    """Qwen3TTS.swift:883-999: [instruct] + role(3) + (pad.. + bos) + codec prefix[:-1], then first text token + last codec embed;
    trailing = text[4:-5] + tts EOS."""
    cfg = small_cfg()
    W = oq.init_weights(cfg, 9, std=0.3)
    t = oq.Talker(cfg, W)
    chat = list(range(10, 10 + 14))                      # 3 role + 1 first + 5 trailing + 5 dropped suffix tokens
    kw = dict(tts_bos=1, tts_eos=2, tts_pad=3, codec_think_id=20, codec_nothink_id=21, codec_think_bos_id=22, codec_think_eos_id=23,
              codec_pad_id=24, codec_bos_id=25)
    x, trailing, pad = oq.prepare_generation_inputs(cfg, W, chat, **kw)
    # no language: prefix [nothink, think_bos, think_eos] + [pad, bos] -> 5 codec embeds, 4 of them under (3 x tts_pad + tts_bos)
    assert x.shape == (1, 3 + 4 + 1, cfg.hidden_size) and trailing.shape == (1, 14 - 9 + 1, cfg.hidden_size)
    text = t.embed_text(torch.tensor([chat]))
    codec = t.embed_codec(torch.tensor([[21, 22, 23, 24, 25]]))
    tts = t.embed_text(torch.tensor([[1, 2, 3]]))
    assert torch.allclose(x[:, :3], text[:, :3])
    assert torch.allclose(x[:, 3:6], tts[:, 2:3] + codec[:, 0:3]) and torch.allclose(x[:, 6:7], tts[:, 0:1] + codec[:, 3:4])
    assert torch.allclose(x[:, 7:8], text[:, 3:4] + codec[:, 4:5])
    assert torch.allclose(trailing[:, :-1], text[:, 4:9]) and torch.allclose(trailing[:, -1:], tts[:, 1:2]) and torch.equal(pad, tts[:, 2:3])
    # language id -> 4-token think prefix; a speaker embedding sits between prefix and [pad, bos]; instruct goes first
    x2, _, _ = oq.prepare_generation_inputs(cfg, W, chat, language_id=30, speaker_id=31, instruct_ids=[40, 41], **kw)
    assert x2.shape[1] == 2 + 3 + (4 + 1 + 2 - 1) + 1
    codec2 = t.embed_codec(torch.tensor([[20, 22, 30, 23, 31, 24, 25]]))
    assert torch.allclose(x2[:, 2 + 3 + 5:2 + 3 + 6], tts[:, 0:1] + codec2[:, 5:6]) and torch.allclose(x2[:, -1:], text[:, 3:4] + codec2[:, 6:7])
    codes = oq.generate_codes(cfg, W, x, trailing, pad, max_tokens=3, temperature=0.0, stop_on_eos=False)
    assert codes.shape == (3, cfg.num_code_groups)


def test_reference_tiny_model_end_to_end_shapes():
    """The geometry of the reference's own Qwen3-TTS tests (Tests/MLXAudioTTSTests.swift:615-687: talker hidden 16 x 2 layers, head_dim 4,
    2 code groups, 1-layer predictor, default-geometry speech-tokenizer decoder) through both oracle halves: prompt embeddings ->
    frame loop (maxTokens 2, T 0.7, top-p 0.95) -> decodeChunk.  The reference asserts audio.ndim == 1 and a non-empty waveform
    (:981-989, :1064-1072)."""
    from oracle import qwen3_tts_codec as oc
    cp = oq.CodePredictorConfig(vocab_size=2048, hidden_size=16, intermediate_size=32, num_hidden_layers=1, num_attention_heads=4,
                                num_key_value_heads=4, head_dim=4, num_code_groups=2)
    cfg = oq.TalkerConfig(vocab_size=3072, hidden_size=16, intermediate_size=32, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=4,
                          head_dim=4, num_code_groups=2, text_hidden_size=16, text_vocab_size=64, codec_eos_token_id=3050, code_predictor=cp)
    W = oq.init_weights(cfg, 3)
    inp, trail, pad = oq.prepare_generation_inputs(cfg, W, chat_ids=[4, 10, 11, 30, 31, 32, 33, 34, 5, 12, 4, 10, 11], tts_bos=22, tts_eos=23, tts_pad=21,
                                                   language_id=3057, codec_think_id=3051, codec_nothink_id=3052, codec_think_bos_id=3053,
                                                   codec_think_eos_id=3054, codec_pad_id=3055, codec_bos_id=3056)
    codes = oq.generate_codes(cfg, W, inp, trail, pad, max_tokens=2, temperature=0.7, top_p=0.95, generator=torch.Generator().manual_seed(0))
    assert codes.shape == (2, 2) and int(codes.min()) >= 0 and int(codes[:, 0].max()) < 3072 - 1024 and int(codes[:, 1].max()) < 2048
    again = oq.generate_codes(cfg, W, inp, trail, pad, max_tokens=2, temperature=0.7, top_p=0.95, generator=torch.Generator().manual_seed(0))
    assert torch.equal(codes, again)                               # same seed, same frames (the reference re-seeds between its two calls)
    dcfg = oc.TokenizerDecoderConfig()                             # "decoder_config": {} in the reference's fixture
    DW = oc.init_weights(dcfg, 1)
    audio = oc.decode_chunk(dcfg, DW, codes[None].numpy())         # [1, frames, groups]: only 2 of the 16 code groups are present
    assert audio.ndim == 1 and audio.shape[0] == 2 * 1920 and np.abs(audio).max() > 0
    parts = oc.streaming_decode(dcfg, DW, codes[None].numpy(), chunk_tokens=1)        # generateStream with a short interval: one chunk per frame
    assert [p.shape for p in parts] == [(1, 1920), (1, 1920)]


def test_filter_logits_matches_transformers_warpers():
    """sampleToken's top-k -> top-p -> min-p chain (Qwen3TTS.swift:1048-1105, applied to the UN-tempered logits) against the independent
    `transformers` warpers that implement the same rules (TopK, TopP with the ascending-cumulative `<= 1 - top_p` removal, MinP relative
    to the most probable token)."""
    from transformers.generation.logits_process import MinPLogitsWarper, TopKLogitsWarper, TopPLogitsWarper
    rng = np.random.default_rng(1)
    for top_k, top_p, min_p in ((50, 1.0, 0.0), (50, 0.9, 0.0), (20, 0.7, 0.05), (0, 0.95, 0.1)):
        for trial in range(3):
            logits = torch.from_numpy((rng.standard_normal((1, 3072)) * 3).astype(np.float32)).double()
            ours = oq.filter_logits(logits, temperature=0.9, top_p=top_p, top_k=top_k, min_p=min_p)
            ref = logits.clone()
            if top_k > 0:
                ref = TopKLogitsWarper(top_k)(None, ref)
            if top_p < 1.0:
                ref = TopPLogitsWarper(top_p)(None, ref)
            if min_p > 0.0:
                ref = MinPLogitsWarper(min_p)(None, ref)
            assert torch.equal(torch.isfinite(ours), torch.isfinite(ref)), (top_k, top_p, min_p, trial)
            keep = torch.isfinite(ref)
            assert torch.allclose(ours[keep], ref[keep])


def test_oracle_reproduces_committed_talker_golden():
    """tests/golden/qwen3_talker.npz (make_golden.py --only qwen3_talker): prompt embeddings, first-step logits and five greedy frames of a
    small talker + code predictor -- the fixture a CUDA talker will be compared with."""
    import importlib.util
    from conftest import GOLDEN
    spec = importlib.util.spec_from_file_location("make_golden", GOLDEN / "make_golden.py")
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    g = np.load(GOLDEN / "qwen3_talker.npz")
    cfg = mg.qwen3_talker_config()
    W = mg.qwen3_talker_weights(cfg)
    inp, trail, pad = oq.prepare_generation_inputs(cfg, W, mg.QWEN3_CHAT_IDS, tts_bos=160, tts_eos=161, tts_pad=162, language_id=2160)
    assert tuple(g["input_shape"]) == tuple(inp.shape) and np.allclose(mg.stats(inp.numpy()), g["input_embeds_stats"], rtol=1e-9, atol=1e-12)
    logits, _ = oq.Talker(cfg, W)(inp, None)
    assert np.array_equal(np.argsort(-logits[0, -1].numpy())[:8], g["first_logits_top"])
    assert np.allclose(mg.stats(logits[0, -1].numpy()), g["first_logits_stats"], rtol=1e-9, atol=1e-12)
    codes = oq.generate_codes(cfg, W, inp, trail, pad, max_tokens=5, temperature=0.0, repetition_penalty=1.05, stop_on_eos=False)
    assert np.array_equal(codes.numpy(), g["codes"])
    assert codes.shape == (5, 4) and int(codes[:, 0].max()) < 3072 - 1024           # the special-token block is suppressed (:383-385)
