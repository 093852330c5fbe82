"""The ALGORITHM of the device sampler (csrc/sampling.cu), modelled in numpy and checked against the pinned oracle
(oracle/sampler_oracle.py ← the reference's own Sampler, tests/test_sampler_oracle.py).

The reference sorts the vocabulary (thrust) up to three times per token.  bf16 logits can take at most 65 536 distinct
values, so the device sampler never sorts: it histograms the order-preserving 16-bit keys of the logits (integer
atomics: exact and deterministic), walks the ≤ 65 536 bins from the largest value down with prefix sums to find the
top-k / top-p / min-p cut — whole bins, plus `m` entries of ONE partially kept bin, which are its lowest indices (the
oracle's tie rule) — and finishes with one pass over the vocabulary in index order (rank inside the partial bin,
probabilities, inclusive cdf, first index with cdf ≥ u·total).  This file restates exactly that plan bin by bin
(`bins_plan`, `bins_sample`) so that the plan — not yet the CUDA — is verified on the CPU, ties included.
"""
import numpy as np
import pytest
import torch

from oracle import sampler_oracle as so

F32 = np.float32


def order_key(bits: np.ndarray) -> np.ndarray:
    """uint16 bf16 bit pattern → uint16 key with the same order as the float value."""
    bits = bits.astype(np.uint16)
    neg = (bits & 0x8000) != 0
    return np.where(neg, ~bits, bits | 0x8000).astype(np.uint16)


def key_value(keys: np.ndarray) -> np.ndarray:
    """inverse of order_key → fp32 value of the bf16 number."""
    keys = keys.astype(np.uint16)
    bits = np.where((keys & 0x8000) != 0, keys & 0x7FFF, ~keys).astype(np.uint16)
    return (bits.astype(np.uint32) << 16).view(np.float32)


def bins_plan(hist, temperature, top_k, top_p, min_p):
    """→ (kept[65536] entries kept per bin, e[65536] = exp(v − vmax) per bin, Z of the kept set).  Bins are walked from
    key 65535 down; every quantity is what one thread-block computes with prefix sums over the bins."""
    hist = hist.astype(np.int64)
    keys = np.arange(65536, dtype=np.uint16)
    v = key_value(keys)
    if temperature > 0:
        with np.errstate(all="ignore"):
            v = (v / F32(temperature)).astype(F32)
    present = hist > 0
    kmax = int(np.max(np.nonzero(present)[0]))
    vmax = v[kmax]
    with np.errstate(all="ignore"):
        e = np.where(present, np.exp((v - vmax).astype(F32)).astype(F32), F32(0)).astype(F32)
    desc = np.arange(65535, -1, -1)                       # walk order
    kept = hist.copy()
    V = int(hist.sum())
    if top_k > 0 and top_k < V:
        cnt_gt = np.zeros(65536, dtype=np.int64)
        cnt_gt[desc] = np.cumsum(hist[desc]) - hist[desc]  # entries with a strictly larger key
        kept = np.clip(top_k - cnt_gt, 0, hist)
    if top_p < 1.0:
        Z = F32(np.sum((kept * e.astype(np.float64))))     # the device sums kept·e in a fixed order; fp64 here
        p = (e / Z).astype(F32)
        mass = kept * p.astype(np.float64)
        cum_before = np.zeros(65536, dtype=np.float64)
        cum_before[desc] = np.cumsum(mass[desc]) - mass[desc]
        with np.errstate(all="ignore"):
            m = np.floor((np.float64(F32(top_p)) - cum_before) / np.where(p > 0, p, 1).astype(np.float64) + 1e-9)
        m = np.clip(m, 0, kept).astype(np.int64)
        m[kmax] = max(m[kmax], 1)                           # the first sorted entry always survives
        kept = np.where(kept > 0, m, 0)
    if min_p > 0:
        Z = F32(np.sum(kept * e.astype(np.float64)))
        p = (e / Z).astype(F32)
        thr = F32(p[kmax] * F32(min_p))
        kept = np.where(p < thr, 0, kept)
    Z = F32(np.sum(kept * e.astype(np.float64)))
    # what sample_draw_kernel relies on: besides whole bins, at most ONE bin is kept partially
    assert int(np.sum((kept > 0) & (kept < hist))) <= 1
    return kept, e, Z


def bins_sample(logits_bf16: torch.Tensor, temperature, top_k, top_p, min_p, u):
    """→ (probs [V] fp32, drawn index) the way the device sampler computes them."""
    bits = logits_bf16.view(torch.int16).numpy().view(np.uint16)
    keys = order_key(bits)
    hist = np.bincount(keys, minlength=65536)
    kept, e, Z = bins_plan(hist, temperature, top_k, top_p, min_p)
    # index-order pass: an element survives if its rank among EQUAL keys (in index order) is below kept[key]
    order = np.argsort(keys, kind="stable")
    rank = np.empty(len(keys), dtype=np.int64)
    sorted_keys = keys[order]
    first = np.searchsorted(sorted_keys, sorted_keys, side="left")
    rank[order] = np.arange(len(keys)) - first
    alive = rank < kept[keys]
    probs = np.where(alive, (e[keys] / Z).astype(F32), F32(0)).astype(F32)
    return probs, so.draw(probs, u)


def _rand_bf16(V, seed, scale):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(V, generator=g) * scale).to(torch.bfloat16)


CFGS = [(0.8, 0, 1.0, 0.0), (1.0, 50, 1.0, 0.0), (0.7, 0, 0.9, 0.0), (1.3, 0, 1.0, 0.05), (0.6, 40, 0.95, 0.02),
        (2.0, 5, 0.5, 0.0), (1.0, 1, 1.0, 0.0), (0.9, 0, 0.0001, 0.0), (1.0, 100000, 0.999, 0.5), (1.0, 7, 0.3, 0.9)]


@pytest.mark.parametrize("V,scale", [(97, 3.0), (5000, 3.0), (151936, 2.0), (4096, 0.05)])
@pytest.mark.parametrize("cfg", CFGS, ids=lambda c: "T{:g}-k{:g}-p{:g}-m{:g}".format(*c))
def test_bin_plan_equals_sort_based_oracle(V, scale, cfg):
    T, k, p, mp = cfg
    logits = _rand_bf16(V, seed=V + int(10 * T) + k, scale=scale)     # bf16: ties are the rule, not the exception
    want = so.filter_probs(logits.float().numpy(), T, k, p, mp)
    got, _ = bins_sample(logits, T, k, p, mp, 0.5)
    differ = np.nonzero((got > 0) != (want > 0))[0]
    if len(differ):
        # only the single top-p boundary entry may flip, and only when the running sum sits on top_p within fp32
        # rounding (the oracle adds probabilities one by one, the bin plan multiplies counts)
        assert len(differ) == 1 and p < 1.0, (cfg, V, differ[:5])
        order = np.argsort(-(logits.float().numpy() / (T if T > 0 else 1)), kind="stable")
        cum = np.cumsum(so.softmax_f32((logits.float().numpy() / F32(T if T > 0 else 1))[order].astype(F32)), dtype=F32)
        pos = int(np.nonzero(order == differ[0])[0][0])
        assert abs(float(cum[pos]) - p) < 2e-5
        return
    np.testing.assert_allclose(got, want, rtol=3e-5, atol=1e-8)
    for u in (0.0003, 0.21, 0.5, 0.77, 0.9996):
        a, b = so.draw(got, u), so.draw(want, u)
        if a != b:  # both cdfs bracket r within rounding
            cdf = np.cumsum(want, dtype=F32)
            assert abs(float(cdf[min(a, b)]) - u * float(cdf[-1])) < 1e-5


def test_key_transform_is_order_preserving_and_invertible():
    bits = np.arange(65536, dtype=np.uint16)
    vals = (bits.astype(np.uint32) << 16).view(np.float32)
    ok = np.isfinite(vals)
    keys = order_key(bits)
    assert len(np.unique(keys)) == 65536
    assert np.array_equal(key_value(keys).view(np.uint32), vals.view(np.uint32))
    idx = np.argsort(keys[ok].astype(np.int64), kind="stable")
    sv = vals[ok][idx]
    assert (np.diff(sv) >= 0).all()      # −0.0 sorts just below +0.0, both compare equal as floats


# --------------------------------------------------------------------------------- thread-level emulation of sampling.cu
PLAN_THREADS, BINS_PER_THREAD = 512, 128
DRAW_THREADS, DRAW_ITEMS = 512, 8


def emulate_plan_kernel(hist, V, temperature, top_k, top_p, min_p):
    """sample_plan_kernel statement by statement: thread t owns bins hi − j (hi = 65535 − 128 t, j = 0…127) and walks
    them downwards; block scans / reductions are exclusive prefix sums / sums over the per-thread values in thread
    order (what cub::BlockScan / BlockReduce deliver)."""
    hist = hist.astype(np.int64)
    T = PLAN_THREADS
    hi = [65535 - t * BINS_PER_THREAD for t in range(T)]
    kept = np.zeros(65536, dtype=np.int64)

    def sv(key):
        v = key_value(np.array([key], dtype=np.uint16))[0]
        return F32(v / F32(temperature)) if temperature > 0 else F32(v)

    my_max = [-1] * T
    for t in range(T):
        for j in range(BINS_PER_THREAD):
            if hist[hi[t] - j] != 0:
                my_max[t] = hi[t] - j
                break
    kmax = max(my_max)
    vmax = sv(kmax)

    def e_of(key):
        with np.errstate(all="ignore"):
            return F32(np.exp(F32(sv(key) - vmax)))

    use_k = 0 < top_k < V
    mine = [int(sum(hist[hi[t] - j] for j in range(BINS_PER_THREAD))) for t in range(T)]
    before = np.concatenate([[0], np.cumsum(mine)[:-1]])
    for t in range(T):
        b = int(before[t])
        for j in range(BINS_PER_THREAD):
            c = int(hist[hi[t] - j])
            k = c
            if use_k:
                k = min(max(top_k - b, 0), c)
            kept[hi[t] - j] = k
            b += c

    def block_z():
        tot = 0.0
        for t in range(T):
            m = 0.0
            for j in range(BINS_PER_THREAD):
                k = int(kept[hi[t] - j])
                if k:
                    m += k * float(e_of(hi[t] - j))
            tot += m
        return tot

    if top_p < 1.0:
        zf = F32(block_z())
        mass = []
        for t in range(T):
            m = 0.0
            for j in range(BINS_PER_THREAD):
                k = int(kept[hi[t] - j])
                if k:
                    m += k * float(F32(e_of(hi[t] - j) / zf))
            mass.append(m)
        before = np.concatenate([[0.0], np.cumsum(mass)[:-1]])
        for t in range(T):
            b = float(before[t])
            for j in range(BINS_PER_THREAD):
                key = hi[t] - j
                k = int(kept[key])
                if k == 0:
                    continue
                p = F32(e_of(key) / zf)
                m = int(np.floor((float(F32(top_p)) - b) / float(p) + 1e-9))
                m = min(max(m, 0), k)
                if key == kmax and m < 1:
                    m = 1
                kept[key] = m
                b += k * float(p)
    if min_p > 0:
        zf = F32(block_z())
        thr = F32(F32(e_of(kmax) / zf) * F32(min_p))
        for key in np.nonzero(kept)[0]:
            if F32(e_of(int(key)) / zf) < thr:
                kept[key] = 0
    z = block_z()
    partial = [(int(k), int(kept[k])) for k in np.nonzero((kept > 0) & (kept < hist))[0]]
    pkey, pkeep = partial[-1] if partial else (-1, 0)
    return kept, z, vmax, pkey, pkeep


def emulate_draw_kernel(keys, kept, z, vmax, pkey, pkeep, temperature, u):
    """sample_draw_kernel: two passes over the vocabulary in chunks of 512 × 8 consecutive entries, carries across
    chunks for the cdf and for the rank inside the partial bin, first index with cdf ≥ u·total."""
    V = len(keys)
    zf = F32(z)
    chunk = DRAW_THREADS * DRAW_ITEMS
    vals = key_value(np.arange(65536, dtype=np.uint16))
    with np.errstate(all="ignore"):
        sv = (vals / F32(temperature)).astype(F32) if temperature > 0 else vals
        e = np.exp((sv - vmax).astype(F32)).astype(F32)
    total, pick = 0.0, -1
    for pas in range(2):
        carry, rank_carry, pick = 0.0, 0, -1
        r = float(F32(u)) * total
        for base in range(0, V, chunk):
            p = np.zeros((DRAW_THREADS, DRAW_ITEMS), dtype=F32)
            isp = np.zeros((DRAW_THREADS, DRAW_ITEMS), dtype=np.int64)
            for t in range(DRAW_THREADS):
                for j in range(DRAW_ITEMS):
                    i = base + t * DRAW_ITEMS + j
                    if i < V and kept[keys[i]] != 0:
                        isp[t, j] = int(keys[i] == pkey)
                        p[t, j] = F32(e[keys[i]] / zf)
            n_partial = isp.sum(axis=1)
            rank_before = np.concatenate([[0], np.cumsum(n_partial)[:-1]])
            mine = np.zeros(DRAW_THREADS)
            last_rank = rank_carry
            for t in range(DRAW_THREADS):
                rank = rank_carry + int(rank_before[t])
                for j in range(DRAW_ITEMS):
                    if isp[t, j]:
                        if rank >= pkeep:
                            p[t, j] = 0
                        rank += 1
                    mine[t] += float(p[t, j])
                if t == DRAW_THREADS - 1:
                    last_rank = rank
            before = np.concatenate([[0.0], np.cumsum(mine)[:-1]])
            chunk_total = float(mine.sum())
            if pas == 1 and pick < 0:
                cands = []
                for t in range(DRAW_THREADS):
                    c = carry + float(before[t])
                    for j in range(DRAW_ITEMS):
                        c += float(p[t, j])
                        i = base + t * DRAW_ITEMS + j
                        if i < V and c >= r:
                            cands.append(i)
                            break
                if cands:
                    pick = min(cands)
            carry += chunk_total
            rank_carry = last_rank
            if pas == 1 and pick >= 0:
                break
        if pas == 0:
            total = carry
    if not total > 0 or pick < 0:
        pick = V - 1 if total > 0 else 0
    return pick


@pytest.mark.parametrize("V,scale,cfg", [(97, 3.0, (0.8, 0, 1.0, 0.0)), (5000, 0.05, (0.6, 40, 0.95, 0.02)),
                                         (9000, 3.0, (1.0, 7, 0.3, 0.9)), (5000, 0.05, (1.0, 50, 1.0, 0.0)),
                                         (4097, 2.0, (0.7, 0, 0.9, 0.0)), (12000, 0.02, (1.3, 3000, 0.97, 0.001))])
def test_thread_level_emulation_of_the_kernels(V, scale, cfg):
    """The per-thread decomposition of sample_plan_kernel / sample_draw_kernel (bins per thread walked downwards,
    exclusive scans over threads, chunk carries, rank inside the partial bin) gives the bin plan's result."""
    T, k, p, mp = cfg
    logits = _rand_bf16(V, seed=V + k, scale=scale)
    bits = logits.view(torch.int16).numpy().view(np.uint16)
    keys = order_key(bits)
    hist = np.bincount(keys, minlength=65536)
    kept_model, e_model, z_model = bins_plan(hist, T, k, p, mp)
    kept, z, vmax, pkey, pkeep = emulate_plan_kernel(hist, V, T, k, p, mp)
    assert np.array_equal(kept, kept_model)
    assert abs(z - float(z_model)) <= 1e-6 * float(z_model)
    for u in (0.0003, 0.21, 0.5, 0.77, 0.9996):
        probs, want = bins_sample(logits, T, k, p, mp, u)
        got = emulate_draw_kernel(keys, kept, z, vmax, pkey, pkeep, T, u)
        if got != want:   # fp64 carries vs the model's fp32 cdf: only at a rounding boundary
            cdf = np.cumsum(probs, dtype=F32)
            assert probs[got] > 0 and abs(float(cdf[min(got, want)]) - u * float(cdf[-1])) < 1e-5, (u, got, want)
