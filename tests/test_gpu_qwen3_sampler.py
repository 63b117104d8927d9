"""The Qwen3-TTS sampler kernel (csrc/qwen3_sampler.cu) through b2a_qwen3_sample_test against the oracle's sampleToken restatement.
Green on the B200 since round 2 (row N1); the same kernel is launched per frame by the talker loop (tests/test_gpu_qwen3_talker.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import qwen3_tts as oq
from qwen3_sampler_model import filtered_row

pytestmark = pytest.mark.gpu


def run(b2a, logits, *, T=0.9, top_p=1.0, top_k=50, min_p=0.0, rep=1.0, eos=-1, suppress=(0, 0), seen=None, track=0, seed=1, step=0, want_filtered=True):
    f = b2a._ffi
    lg = np.ascontiguousarray(logits, np.float32)
    B, V = lg.shape
    words = (V + 31) // 32
    bitmap = None
    if seen is not None:
        bitmap = np.zeros((B, words), np.uint32)
        for b, toks in enumerate(seen):
            for t in toks:
                bitmap[b, t >> 5] |= np.uint32(1 << (t & 31))
    toks = np.zeros(B, np.int32)
    filt = np.zeros((B, V), np.float32) if want_filtered else None
    f.check(f.lib().b2a_qwen3_sample_test(f.ptr(lg), B, V, T, top_p, top_k, min_p, rep, eos, suppress[0], suppress[1], f.ptr(bitmap), track, seed, step,
                                          f.ptr(toks), f.ptr(filt)))
    return toks, filt, bitmap


@pytest.mark.parametrize("V,top_k,top_p,min_p,rep,T", [(3072, 50, 1.0, 0.0, 1.0, 0.9), (3072, 50, 0.95, 0.0, 1.05, 0.7), (2048, 50, 0.8, 0.0, 1.0, 0.9),
                                                         (2048, 0, 0.9, 0.05, 1.0, 1.0), (97, 200, 0.5, 0.0, 1.3, 0.6)])
def test_filtered_logits_match_oracle(b2a, V, top_k, top_p, min_p, rep, T):
    rng = np.random.default_rng(V + top_k)
    B = 5
    logits = (rng.standard_normal((B, V)) * 3.0).astype(np.float32)
    eos = V - 1024 + 2 if V > 1100 else -1
    suppress = (V - 1024, V) if V > 1100 else (0, 0)
    seen = [rng.integers(0, V, size=17).tolist() if rep != 1.0 else [] for _ in range(B)]
    toks, filt, _ = run(b2a, logits, T=T, top_p=top_p, top_k=top_k, min_p=min_p, rep=rep, eos=eos, suppress=suppress, seen=seen)
    for b in range(B):
        ref = oq.filter_logits(torch.from_numpy(logits[b])[None], temperature=T, top_p=top_p, top_k=top_k, repetition_penalty=rep, generated_tokens=seen[b],
                               suppress_tokens=[t for t in range(*suppress) if t != eos], eos_token_id=eos if eos >= 0 else None, min_p=min_p)[0].numpy()
        keep = np.isfinite(ref)
        assert np.array_equal(np.isfinite(filt[b]), keep), (b, np.flatnonzero(np.isfinite(filt[b]) != keep))
        assert np.abs(filt[b][keep] - ref[keep]).max() < 1e-5
        assert keep[toks[b]]                                       # the drawn token is one the filters left in
        model = filtered_row(logits[b], temperature=T, top_p=top_p, top_k=top_k, min_p=min_p, rep_penalty=rep, eos=eos, suppress=suppress, seen=seen[b])
        assert np.array_equal(np.isfinite(filt[b]), np.isfinite(model))      # and the step-for-step numpy model of the kernel


def test_greedy_and_seen_bitmap(b2a):
    rng = np.random.default_rng(0)
    logits = (rng.standard_normal((4, 3072)) * 2.0).astype(np.float32)
    logits[1, 77] = logits[1, 900] = logits[1].max() + 1.0          # a tie: the lower index wins (MLX argMax)
    seen = [[5, 6], [], [int(np.argmax(logits[2, :2048]))], []]
    toks, filt, bitmap = run(b2a, logits, T=0.0, rep=1.5, eos=2050, suppress=(2048, 3072), seen=seen, track=1)
    for b in range(4):
        ref = oq.filter_logits(torch.from_numpy(logits[b])[None], temperature=0.0, repetition_penalty=1.5, generated_tokens=seen[b],
                               suppress_tokens=[t for t in range(2048, 3072) if t != 2050], eos_token_id=2050)[0]
        assert toks[b] == int(ref.argmax()), b
        assert (bitmap[b, toks[b] >> 5] >> (toks[b] & 31)) & 1       # the kernel recorded its own token
    assert toks[1] == 77


def test_draw_frequencies_follow_the_filtered_distribution(b2a):
    rng = np.random.default_rng(3)
    V, T = 2048, 0.8
    row = (rng.standard_normal(V) * 2.0).astype(np.float32)
    logits = np.repeat(row[None], 512, axis=0)                       # 512 rows = 512 independent draws (the row index feeds the RNG)
    counts = np.zeros(V)
    for step in range(8):
        toks, _, _ = run(b2a, logits, T=T, top_k=8, want_filtered=False, seed=11, step=step)
        np.add.at(counts, toks, 1)
    f = oq.filter_logits(torch.from_numpy(row)[None], temperature=T, top_k=8)[0]
    p = torch.softmax(f / T, dim=-1).numpy()
    assert counts[p == 0].sum() == 0
    n = counts.sum()
    big = p > 0.02
    assert np.abs(counts[big] / n - p[big]).max() < 4.0 * np.sqrt(p[big].max() / n)
