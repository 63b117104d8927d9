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
