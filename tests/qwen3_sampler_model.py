"""numpy model of q3s::sample_kernel (csrc/qwen3_sampler.cu), step for step: one descending sort, top-k as a prefix, top-p as
suffix sums, min-p, EOS write-back.  Shared by the CPU design test and the gated GPU test.  Test infrastructure."""
import numpy as np


def filtered_row(logits, *, temperature=0.9, top_p=1.0, top_k=50, min_p=0.0, rep_penalty=1.0, eos=-1, suppress=(0, 0), seen=()):
    x = np.asarray(logits, dtype=np.float32).copy()
    V = x.shape[0]
    lo, hi = suppress
    for i in range(max(lo, 0), min(hi, V)):
        if i != eos:
            x[i] = -np.inf
    if rep_penalty != 1.0:
        for tkn in set(seen):
            if tkn < V:
                x[tkn] = x[tkn] * rep_penalty if x[tkn] < 0 else x[tkn] / rep_penalty
    if temperature <= 0:
        return x
    eos_logit = x[eos] if 0 <= eos < V else None
    order = np.lexsort((np.arange(V), -x))                      # (logit desc, index asc): the kernel's `before`
    key = x[order].copy()
    if 0 < top_k < V:
        key[top_k:] = -np.inf
    top = key[0]
    if 0.0 < top_p < 1.0:
        e = np.exp(key - top, dtype=np.float32)
        suffix = np.cumsum(e[::-1], dtype=np.float32)[::-1]
        key[~(suffix > np.float32(1.0 - top_p) * suffix[0])] = -np.inf
    if min_p > 0.0:
        key[key < top + np.float32(np.log(min_p))] = -np.inf
    out = np.full(V, -np.inf, np.float32)
    out[order] = key
    if eos_logit is not None:
        out[eos] = eos_logit
    return out
