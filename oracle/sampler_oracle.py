"""TEST INFRASTRUCTURE — CPU restatement of the reference's sampler (the part of the async token
pipeline that is not greedy).  Nothing in the product imports this.

    tinygpt::Sampler::sample           src/engine/Sampler.cpp:23-78
    multinomial (nSamples = 1)         third_party/TinyTorch/src/Operation/OpSamplingCuda.cu:30-62,230-247

Pipeline on logits [V] (fp32 here; the reference runs the same ops in the tensor's dtype):
    greedy when temperature ≤ 0, top_k ≤ 0, top_p ≥ 1, min_p ≤ 0            (Sampler.cpp:17-29: argmax, LAST index wins on
                                                                             the CUDA path — decode_oracle.argmax_last)
    l ← l / temperature                         if temperature > 0          (:34-36)
    keep the top_k largest, others ← −inf       if top_k > 0                (:39-45)
    sort descending, softmax, inclusive cumsum; keep entries whose cumsum ≤ top_p, and ALWAYS the first
                                                if top_p < 1                (:48-65)
    p ← softmax(l); drop entries with p < max(p)·min_p   if min_p > 0       (:68-74)
    probs ← softmax(l)                                                      (:77)
    draw: cdf = inclusive cumsum of probs in INDEX order, r = u·cdf[-1], first index with cdf ≥ r   (multinomial)

Ties.  The reference sorts with thrust (CUDA: stable_sort_by_key for `sort`, unstable sort_by_key for `topk`) or
std::sort / partial_sort (CPU), so which of several EQUAL logits survives a top-k / top-p boundary is unspecified
there; here equal values are ordered by ascending index (what a stable descending sort gives).  With continuous random
logits ties have probability zero, which is what the fixtures use; bf16 logits do tie.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def softmax_f32(x: np.ndarray) -> np.ndarray:
    m = np.max(x)
    e = np.exp((x - m).astype(F32)).astype(F32)
    e[np.isneginf(x)] = 0.0
    return (e / np.sum(e, dtype=F32)).astype(F32)


def filter_probs(logits: np.ndarray, temperature: float = 0.0, top_k: int = 0, top_p: float = 1.0,
                 min_p: float = 0.0) -> np.ndarray:
    """Probabilities the reference's sampler hands to multinomial (fp32 [V])."""
    l = np.asarray(logits, dtype=F32).copy()
    V = l.shape[0]
    if temperature > 0:
        l = (l / F32(temperature)).astype(F32)
    if top_k > 0:
        k = min(int(top_k), V)
        order = np.argsort(-l, kind="stable")          # descending, ties by ascending index
        keep = order[:k]
        out = np.full(V, -np.inf, dtype=F32)
        out[keep] = l[keep]
        l = out
    if top_p < 1.0:
        order = np.argsort(-l, kind="stable")
        sl = l[order]
        probs = softmax_f32(sl)
        cum = np.cumsum(probs, dtype=F32)
        mask = cum <= F32(top_p)
        mask[0] = True                                  # firstMask (:54-58)
        sl = np.where(mask, sl, F32(-np.inf)).astype(F32)
        out = np.full(V, -np.inf, dtype=F32)
        out[order] = sl
        l = out
    if min_p > 0:
        p = softmax_f32(l)
        thr = F32(np.max(p) * F32(min_p))
        l = np.where(p < thr, F32(-np.inf), l).astype(F32)
    return softmax_f32(l)


def is_greedy(temperature: float, top_k: int, top_p: float, min_p: float) -> bool:
    return not (temperature > 0 or top_k > 0 or top_p < 1.0 or min_p > 0)


def draw(probs: np.ndarray, u: float) -> int:
    """kMultinomialWithReplacement: first index whose inclusive fp32 cdf (index order) reaches r = u·total."""
    cdf = np.cumsum(np.asarray(probs, dtype=F32), dtype=F32)
    total = cdf[-1]
    if not total > 0:
        return 0
    r = F32(F32(u) * total)
    return int(np.searchsorted(cdf, r, side="left"))


def sample(logits: np.ndarray, temperature: float, top_k: int, top_p: float, min_p: float, u: float) -> int:
    if is_greedy(temperature, top_k, top_p, min_p):
        l = np.asarray(logits, dtype=F32)
        return int(len(l) - 1 - np.argmax(l[::-1]))      # last index among equal maxima (CUDA path)
    return draw(filter_probs(logits, temperature, top_k, top_p, min_p), u)
