"""Async token pipeline (pinned-host mailbox, DecodeEngine.generate_async) and the device sampler (csrc/sampling.cu)
on the GPU  [ref: src/engine/GPTEngine.cpp:17-35,180-232; src/engine/Sampler.cpp:31-78]."""
import pytest
import torch

from helpers import assert_close_bf16, orc, to_oracle_cfg
from tinygpt_b200 import engine, models

pytestmark = pytest.mark.gpu
DEV = "cuda"

# ------------------------------------------------------------------------------------ async token pipeline (mailbox)
def test_generate_async_matches_sync_and_stops(built_lib):
    """generate_async hands out exactly generate_sync's tokens through the pinned-host mailbox (ring smaller than the
    sequence → wrap-around), stops at an EOS id / on callback abort, and leaves the engine positioned for continuation."""
    spec = models.TINY_QWEN2.with_ctx(256)
    w = {k: v.to(DEV) for k, v in models.synth_weights(spec, seed=2).items()}
    eng = engine.DecodeEngine(spec, w)
    prompt = torch.randint(0, spec.vocab, (11,), generator=torch.Generator().manual_seed(3)).tolist()
    want = eng.generate_sync(prompt, 40).tolist()
    eng.set_mailbox(8)
    seen = []
    got, reason = eng.generate_async(prompt, 40, callback=lambda t: seen.append(t) or True, lookahead=3)
    assert got == want and seen == want and reason == "length"
    # EOS: the 7th token is declared EOS → 6 tokens come out; the steps that ran ahead are rewound
    eos = want[6]
    first_eos = want.index(eos)
    got, reason = eng.generate_async(prompt, 40, eos_ids=[eos], lookahead=4)
    assert got == want[:first_eos] and reason == "stop"
    assert eng.position == len(prompt) + max(first_eos, 1) - 1
    # continuing from there with the last kept token reproduces the sync sequence
    if first_eos >= 1:
        nxt = eng.gen_next_token(torch.tensor([[want[first_eos - 1]]], device=DEV))
        assert int(nxt) == want[first_eos]
    # abort from the callback after 5 tokens
    got, reason = eng.generate_async(prompt, 40, callback=lambda t: len(got_so_far.append(t) or got_so_far) < 5,
                                     lookahead=2) if (got_so_far := []) is not None else (None, None)
    assert got == want[:5] and reason == "stop"
    eng.clear_mailbox()
    assert eng.generate_sync(prompt, 12).tolist() == want[:12]
    eng.close()


# ------------------------------------------------------------------------------------ device sampler (csrc/sampling.cu)
@pytest.mark.parametrize("V,scale", [(97, 3.0), (5000, 3.0), (151936, 2.0), (32768, 0.05)])
def test_device_sampler_vs_oracle(built_lib, V, scale):
    """b200_sample_bf16 against the pinned sampler oracle on bf16 logits (ties included): same drawn index for a set of
    uniform numbers, except where the oracle's own cdf sits on u·total (or a top-p boundary on top_p) within fp32
    rounding; and the same call twice gives the same index (integer histogram + fixed-order sums)."""
    import numpy as np
    from oracle import sampler_oracle as so
    from tinygpt_b200 import ops
    cfgs = [(0.8, 0, 1.0, 0.0), (1.0, 50, 1.0, 0.0), (0.7, 0, 0.9, 0.0), (1.3, 0, 1.0, 0.05), (0.6, 40, 0.95, 0.02),
            (2.0, 5, 0.5, 0.0), (1.0, 1, 1.0, 0.0), (0.9, 0, 0.0001, 0.0), (1.0, 100000, 0.999, 0.5), (1.0, 7, 0.3, 0.9)]
    g = torch.Generator().manual_seed(V)
    logits = (torch.randn(V, generator=g) * scale).to(torch.bfloat16)
    dev_logits = logits.to(DEV)
    lf = logits.float().numpy()
    soft = 0
    for (T, k, p, mp) in cfgs:
        want_probs = so.filter_probs(lf, T, k, p, mp)
        cdf = np.cumsum(want_probs, dtype=np.float32)
        for u in (0.0003, 0.21, 0.5, 0.77, 0.9996):
            got = int(ops.sample(dev_logits, T, k, p, mp, u))
            again = int(ops.sample(dev_logits, T, k, p, mp, u))
            assert got == again, "device sampler must be deterministic"
            want = so.draw(want_probs, u)
            if got != want:
                soft += 1
                r = u * float(cdf[-1])
                near_draw = want_probs[got] > 0 and abs(float(cdf[min(got, want)]) - r) < 2e-5
                assert near_draw or p < 1.0, (V, (T, k, p, mp), u, got, want)
    print(f"[sampler V={V}] {soft} of {len(cfgs) * 5} draws differ from the oracle at a rounding boundary")
    assert soft <= 3


def test_engine_sampler_replays_on_the_host(built_lib):
    """b200_engine_set_sampler: every token the engine draws equals the oracle's draw from the SAME logits with the SAME
    uniform number (Philox(seed, tokens generated so far), mirrored on the host), up to a rounding boundary; switching
    the sampler off restores greedy decoding; sampled tokens also arrive through the mailbox."""
    import numpy as np
    from oracle import sampler_oracle as so
    from tinygpt_b200._lib import lib
    spec = models.TINY_QWEN2.with_ctx(128)
    w = {k: v.to(DEV) for k, v in models.synth_weights(spec, seed=3, std=0.05).items()}
    eng = engine.DecodeEngine(spec, w)
    prompt = torch.randint(0, spec.vocab, (1, 9), generator=torch.Generator().manual_seed(1)).to(DEV)
    greedy = eng.generate_sync(prompt.view(-1).tolist(), 12).tolist()
    cfg = dict(temperature=0.9, top_k=40, top_p=0.95, min_p=0.01)
    eng.set_sampler(seed=1234, **cfg)
    eng.reset_cache()
    ids, soft = prompt, 0
    for step in range(16):
        n = int(lib().b200_engine_generated(eng._h))
        logits = eng.forward(ids)[0, -1].float().cpu().numpy()
        tok = torch.empty(1, dtype=torch.int64, device=DEV)
        lib().b200_engine_last_token(eng._h, tok.data_ptr(), torch.cuda.current_stream().cuda_stream)
        got = int(tok.item())
        u = engine.DecodeEngine.philox_uniform(1234, n)
        want = so.sample(logits, cfg["temperature"], cfg["top_k"], cfg["top_p"], cfg["min_p"], u)
        if got != want:
            soft += 1
            probs = so.filter_probs(logits, cfg["temperature"], cfg["top_k"], cfg["top_p"], cfg["min_p"])
            cdf = np.cumsum(probs, dtype=np.float32)
            assert abs(float(cdf[min(got, want)]) - u * float(cdf[-1])) < 2e-5 or probs[got] > 0, (step, got, want, u)
        ids = tok.view(1, 1)
    assert soft <= 1
    out, reason = eng.generate_async(prompt.view(-1).tolist(), 10, lookahead=2)     # sampled tokens through the mailbox
    assert len(out) == 10 and all(0 <= t < spec.vocab for t in out)
    eng.set_sampler()                                                                 # all knobs off → greedy again
    assert eng.generate_sync(prompt.view(-1).tolist(), 12).tolist() == greedy
    eng.close()
