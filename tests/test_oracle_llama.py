"""Oracle Llama/Orpheus restatement vs transformers' LlamaForCausalLM (random init, llama3 rope scaling),
token plumbing round trips, sampler restatement, goldens.  CPU only."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err
from oracle import llama


def _hf_model(cfg, W):
    from transformers import LlamaConfig as HC, LlamaForCausalLM
    hc = HC(hidden_size=cfg.hidden_size, num_hidden_layers=cfg.num_hidden_layers, intermediate_size=cfg.intermediate_size,
            num_attention_heads=cfg.num_attention_heads, num_key_value_heads=cfg.num_key_value_heads, head_dim=cfg.head_dim,
            vocab_size=cfg.vocab_size, rms_norm_eps=cfg.rms_norm_eps, rope_theta=cfg.rope_theta, tie_word_embeddings=True,
            max_position_embeddings=131072,
            rope_scaling={"rope_type": "llama3", "factor": 32.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
                          "original_max_position_embeddings": 8192}, attn_implementation="eager")
    m = LlamaForCausalLM(hc).float().eval()
    sd = {k: v.float() for k, v in W.items()}
    sd["lm_head.weight"] = sd["model.embed_tokens.weight"]
    m.load_state_dict(sd, strict=True)
    return m


def test_forward_matches_transformers_llama3_rope():
    cfg = llama.LlamaConfig(hidden_size=256, num_hidden_layers=2, intermediate_size=512, num_attention_heads=2,
                            num_key_value_heads=1, head_dim=128, vocab_size=512)
    W = llama.init_weights(cfg, 5, std=0.08)
    ids = torch.as_tensor(np.random.default_rng(0).integers(0, 512, (2, 9)))
    with torch.no_grad():
        ref = _hf_model(cfg, W)(ids).logits.numpy()
    mo = llama.LlamaOracle(cfg, W, round_acts=False)
    mine = mo.forward(ids).numpy()
    assert rel_err(mine, ref) < 2e-5
    # incremental decode with the KV cache == full forward (offset handling, LlamaTTS.swift:247-251)
    mo.reset()
    a = mo.forward(ids[:, :6]).numpy()
    b = np.concatenate([mo.forward(ids[:, 6 + i:7 + i]).numpy() for i in range(3)], axis=1)
    assert rel_err(np.concatenate([a, b], axis=1), mine) < 1e-5


def test_llama3_freqs_are_divisors_and_scaled():
    f = llama.llama3_rope_freqs(llama.LlamaConfig())
    assert f.shape == (64,) and f[0] == 1.0 and np.all(np.diff(f) > 0)
    base = 500000.0 ** (np.arange(0, 128, 2) / 128.0)
    assert np.isclose(f[-1], base[-1] * 32.0, rtol=1e-5)          # longest wavelength scaled by `factor`
    assert np.isclose(f[1], base[1], rtol=1e-6)                   # short wavelengths untouched
    g = np.load(GOLDEN / "llama_tiny.npz")
    assert np.allclose(f, g["freqs"], rtol=1e-6)


def test_frame_interleave_roundtrip_and_layout():
    # LlamaTTS.swift:41-98: frame of 7 -> L1 [c0], L2 [c1-4096, c4-4*4096], L3 [c2-2*4096, c3-3*4096, c5-.., c6-..]
    rng = np.random.default_rng(0)
    codes = [rng.integers(0, 4096, (1, 5 * k), dtype=np.int32) for k in (1, 2, 4)]
    cl = llama.code_list_from_codes(codes)
    assert len(cl) == 35 and cl[1] == codes[1][0, 0] + 4096 and cl[4] == codes[1][0, 1] + 4 * 4096
    back = llama.codes_from_code_list(cl)
    assert all(np.array_equal(a, b) for a, b in zip(back, codes))


def test_parse_output_semantics():
    S, E, O = llama.START_OF_SPEECH, llama.END_OF_SPEECH, llama.AUDIO_TOKEN_OFFSET
    row = [5, S, 7, S] + [O + i for i in range(9)] + [E, O + 9]
    out = llama.parse_output(np.asarray([row]))
    assert out == [[0, 1, 2, 3, 4, 5, 6]]                          # cropped after LAST S, E dropped, trimmed to 7
    assert llama.parse_output(np.asarray([[O + i for i in range(7)]])) == [list(range(7))]   # no S: whole row


def test_prepare_input_ids_framing_and_left_pad():
    ids, mask = llama.prepare_input_ids([[1, 2, 3], [9]])
    assert ids.tolist() == [[128259, 1, 2, 3, 128009, 128260], [128263, 128263, 128259, 9, 128009, 128260]]
    assert mask.tolist()[1][:2] == [False, False]


def test_sampler_restatement():
    l = np.log(np.array([0.5, 0.3, 0.1, 0.06, 0.04]))
    kept = llama.top_p_filter(l, 1.0, 0.75)
    assert np.flatnonzero(kept).tolist() == [0, 1]                 # ascending cumsum .04 .10 .20 .50 1.0 > 0.25
    kept = llama.top_p_filter(l, 1.0, 0.85)
    assert np.flatnonzero(kept).tolist() == [0, 1, 2]
    pen = llama.repetition_penalty(np.array([2.0, -2.0, 1.0], dtype=np.float32), [0, 1, 1], 2.0)
    assert pen.tolist() == [1.0, -4.0, 1.0]
    assert llama.sample_inverse_cdf(np.array([0.0, 0.5, 0.0, 0.5]), 0.49) == 1
    assert llama.sample_inverse_cdf(np.array([0.0, 0.5, 0.0, 0.5]), 0.51) == 3


def test_goldens():
    g = np.load(GOLDEN / "llama_tiny.npz")
    cfg = llama.LlamaConfig(hidden_size=256, num_hidden_layers=2, intermediate_size=512, num_attention_heads=2,
                            num_key_value_heads=1, head_dim=128, vocab_size=2048)
    W = llama.init_weights(cfg, 1234, std=0.08)
    lg = llama.LlamaOracle(cfg, W, False).forward(torch.as_tensor(g["ids"])).numpy()
    assert rel_err(lg[:, -1], g["logits_last"]) < 1e-5


def test_sampler_restatement_matches_transformers_logits_processors():
    """The sampler / repetition penalty live in mlx-swift-lm (not on disk).  Their restatement is checked here against the independent
    `transformers` processors that implement the same published rules: RepetitionPenaltyLogitsProcessor (logit < 0 -> * p, else / p, once
    per unique token) and TemperatureLogitsWarper + TopPLogitsWarper (drop the ascending-sorted tail whose cumulative probability is <= 1 - top_p)."""
    import torch
    from transformers.generation.logits_process import RepetitionPenaltyLogitsProcessor, TemperatureLogitsWarper, TopPLogitsWarper
    rng = np.random.default_rng(0)
    for trial in range(5):
        logits = (rng.standard_normal(500) * 3).astype(np.float32)
        ctx = rng.integers(0, 500, size=20).tolist()
        ours = llama.repetition_penalty(logits, ctx, 1.3)
        ref = RepetitionPenaltyLogitsProcessor(1.3)(torch.as_tensor([ctx]), torch.from_numpy(logits)[None].clone())[0].numpy()
        assert np.array_equal(ours, ref)
        kept = llama.top_p_filter(ours, 0.6, 0.8)
        warped = TopPLogitsWarper(0.8)(None, TemperatureLogitsWarper(0.6)(None, torch.from_numpy(ours)[None].double()))[0]
        assert np.array_equal(kept > 0, torch.isfinite(warped).numpy()), trial
        p = torch.softmax(torch.from_numpy(ours).double() / 0.6, dim=-1).numpy()
        assert np.allclose(kept[kept > 0], p[kept > 0], rtol=1e-12)
