"""oracle/sampler_oracle.py against the reference's OWN sampler: tests/golden/ref_cpu_sampler.npz holds what
tinygpt::Sampler::sample (src/engine/Sampler.cpp, compiled unmodified into oracle/_ref) handed to `multinomial` for ten
configurations (temperature / top-k / top-p / min-p, alone and combined, the degenerate ones included) and the index it
drew for four uniform numbers through the hooked multinomial (tests/golden/make_ref_fixtures.py).  The restatement
must give the same support, the same probabilities (fp32 summation-order tolerance) and the same drawn indices."""
from pathlib import Path

import numpy as np
import pytest

from oracle import sampler_oracle as so

FIX = np.load(Path(__file__).resolve().parent / "golden" / "ref_cpu_sampler.npz")
CFGS = [tuple(r) for r in FIX["cfgs"]]


@pytest.mark.parametrize("V", [97, 2048])
@pytest.mark.parametrize("ci", range(len(CFGS)), ids=lambda i: "T{:g}-k{:g}-p{:g}-m{:g}".format(*CFGS[i]))
def test_filtering_and_draw_match_the_reference(V, ci):
    T, k, p, mp = CFGS[ci]
    k = int(k)
    logits, ref_probs, ref_picks = FIX[f"logits_{V}"], FIX[f"probs_{V}"][ci], FIX[f"picks_{V}"][ci]
    if so.is_greedy(T, k, p, mp):
        # greedy branch: argmax, no multinomial (captured probabilities stay zero); continuous logits have no ties, so
        # the CPU reference's first-max and the CUDA path's last-max coincide
        assert not ref_probs.any()
        assert all(int(x) == so.sample(logits, T, k, p, mp, 0.5) for x in ref_picks)
        return
    mine = so.filter_probs(logits, T, k, p, mp)
    assert ((mine > 0) == (ref_probs > 0)).all(), "different surviving set"
    np.testing.assert_allclose(mine, ref_probs, rtol=2e-5, atol=1e-7)
    assert abs(float(mine.sum()) - 1.0) < 1e-5
    for u, want in zip(FIX["u"], ref_picks):
        assert so.draw(mine, float(u)) == int(want)
        assert so.sample(logits, T, k, p, mp, float(u)) == int(want)


def test_tie_rule_and_edge_cases():
    # equal logits: ascending index order inside a tie group (what a stable descending sort gives)
    l = np.array([1.0, 3.0, 3.0, 3.0, 0.0], dtype=np.float32)
    p = so.filter_probs(l, 1.0, top_k=2)
    assert (p > 0).tolist() == [False, True, True, False, False]
    # top-p keeps at least the first sorted entry, however small top_p is
    p = so.filter_probs(l, 1.0, top_p=1e-6)
    assert (p > 0).tolist() == [False, True, False, False, False] and p[1] == 1.0
    # min-p relative to the maximum probability
    l = np.log(np.array([0.5, 0.3, 0.15, 0.05], dtype=np.float32))
    assert (so.filter_probs(l, 1.0, min_p=0.5) > 0).tolist() == [True, True, False, False]
    # u at the ends of the interval
    p = np.array([0.0, 0.25, 0.0, 0.75], dtype=np.float32)
    assert so.draw(p, 0.0) == 0 and so.draw(p, 1e-9) == 1 and so.draw(p, 0.25) == 1 and so.draw(p, 0.2500001) == 3
    assert so.draw(np.zeros(4, dtype=np.float32), 0.3) == 0
    # greedy with ties: last index (CUDA argmax rule)
    assert so.sample(np.array([2.0, 5.0, 5.0, 1.0], dtype=np.float32), 0.0, 0, 1.0, 0.0, 0.5) == 2
