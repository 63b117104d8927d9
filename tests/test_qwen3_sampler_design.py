"""The Qwen3-TTS sampler kernel's ALGORITHM (one descending sort; top-k = prefix, top-p = suffix sums, min-p, EOS write-back), as a
numpy model, against the oracle's restatement of sampleToken (oracle/qwen3_tts.filter_logits).  The CUDA kernel itself has not run
on a GPU yet (tests/test_gpu_qwen3_sampler.py is gated); this checks that what it was written to compute is the right thing.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import qwen3_tts as oq
from qwen3_sampler_model import filtered_row


@pytest.mark.parametrize("V,top_k,top_p,min_p,rep,T", [(3072, 50, 1.0, 0.0, 1.0, 0.9), (3072, 50, 0.95, 0.0, 1.05, 0.7), (2048, 50, 0.8, 0.0, 1.0, 0.9),
                                                         (2048, 0, 0.9, 0.05, 1.0, 1.0), (97, 200, 0.5, 0.0, 1.3, 0.6), (3072, 50, 0.95, 0.02, 1.1, 0.0)])
def test_model_matches_oracle_filter_logits(V, top_k, top_p, min_p, rep, T):
    rng = np.random.default_rng(V + top_k)
    for trial in range(6):
        logits = (rng.standard_normal(V) * 3.0).astype(np.float32)
        eos = V - 1024 + 2 if V > 1100 else -1
        suppress = (V - 1024, V) if V > 1100 else (0, 0)
        seen = rng.integers(0, V, size=17).tolist() if rep != 1.0 else []
        ref = oq.filter_logits(torch.from_numpy(logits)[None], temperature=T, top_p=top_p, top_k=top_k, repetition_penalty=rep, generated_tokens=seen,
                               suppress_tokens=[t for t in range(*suppress) if t != eos], eos_token_id=eos if eos >= 0 else None, min_p=min_p)[0].numpy()
        got = filtered_row(logits, temperature=T, top_p=top_p, top_k=top_k, min_p=min_p, rep_penalty=rep, eos=eos, suppress=suppress, seen=seen)
        assert np.array_equal(np.isfinite(got), np.isfinite(ref)), (trial, np.flatnonzero(np.isfinite(got) != np.isfinite(ref)))
        keep = np.isfinite(ref)
        assert np.abs(got[keep] - ref[keep]).max() < 1e-5
        if T > 0 and eos >= 0:
            assert np.isfinite(got[eos])                          # the EOS logit always survives the filters (Qwen3TTS.swift:1041-1046, 1107-1110)


def test_bitonic_network_as_written_sorts_by_logit_then_index():
    """The compare-exchange schedule of q3s::sample_kernel (k = 2..4096, j = k/2..1, partner i ^ j, direction from i & k) with its
    `before` comparator, emulated slot by slot: descending logits, ties to the lower index, -inf padding at the end."""
    SLOTS, V = 4096, 3072
    rng = np.random.default_rng(1)
    vals = np.round(rng.standard_normal(V) * 3, 1)                 # rounding creates ties
    key = np.full(SLOTS, -np.inf); key[:V] = vals
    idx = np.full(SLOTS, 0x7fffffff, dtype=np.int64); idx[:V] = np.arange(V)
    before = lambda ka, ia, kb, ib: (ka > kb) | ((ka == kb) & (ia < ib))
    k = 2
    while k <= SLOTS:
        j = k >> 1
        while j > 0:
            i = np.arange(SLOTS); p = i ^ j; m = p > i
            ii, pp = i[m], p[m]
            ka, kb, ia, ib = key[ii], key[pp], idx[ii], idx[pp]
            sw = np.where((ii & k) == 0, before(kb, ib, ka, ia), before(ka, ia, kb, ib))
            key[ii], key[pp] = np.where(sw, kb, ka), np.where(sw, ka, kb)
            idx[ii], idx[pp] = np.where(sw, ib, ia), np.where(sw, ia, ib)
            j >>= 1
        k <<= 1
    assert np.array_equal(idx[:V], np.lexsort((np.arange(V), -vals))) and np.all(idx[V:] == 0x7fffffff)


@pytest.mark.parametrize("rev", [False, True])
def test_block_scan_as_written(rev):
    """q3s::block_scan emulated thread by thread (4 slots per thread, warp shuffles with the lane guards, warp totals, exclusive base):
    inclusive prefix sums forward, suffix sums in reverse."""
    PER, THREADS = 4, 1024
    x = np.random.default_rng(0).random(PER * THREADS)
    v = x.reshape(THREADS, PER).copy()
    loc = np.zeros(THREADS)
    for i in (range(PER) if not rev else range(PER - 1, -1, -1)):
        loc += v[:, i]; v[:, i] = loc
    inc = loc.copy()
    lane, warp = np.arange(THREADS) % 32, np.arange(THREADS) // 32
    o = 1
    while o < 32:
        src = np.arange(THREADS) + (o if rev else -o)
        ok = (src >= 0) & (src < THREADS) & ((np.clip(src, 0, THREADS - 1) // 32) == warp)
        n = np.where(ok, inc[np.clip(src, 0, THREADS - 1)], inc)   # an out-of-range source lane returns the caller's own value
        inc = np.where((lane + o < 32) if rev else (lane >= o), inc + n, inc)
        o <<= 1
    wsum = inc[lane == (0 if rev else 31)]
    base = np.array([wsum[(np.arange(32) > w) if rev else (np.arange(32) < w)].sum() for w in warp])
    out = (v + (base + inc - loc)[:, None]).reshape(-1)
    ref = np.cumsum(x[::-1])[::-1] if rev else np.cumsum(x)
    assert np.abs(out - ref).max() < 1e-9 and abs(wsum.sum() - x.sum()) < 1e-9
