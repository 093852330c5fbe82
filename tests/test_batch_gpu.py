"""Batched decode (B ≤ 8 sequences at the same position — GPTEngine::generateSync's left-padded batch,
src/engine/GPTEngine.cpp:101-174): one weight pass per step for the whole batch (csrc/gemv_batch.cu), attention with
grid.z = B, per-sequence KV caches.  Per sequence the arithmetic is the batch-1 kernels', so a batched run must
reproduce B independent batch-1 runs BIT FOR BIT — logits of every step and greedy ids."""
import pytest
import torch

from tinygpt_b200 import engine, models

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _single_runs(spec, w, prompts, n_new):
    """B independent batch-1 runs on a fresh engine each: logits of the prompt + teacher-forced steps, and free ids."""
    out_logits, out_toks = [], []
    for b in range(prompts.shape[0]):
        eng = engine.DecodeEngine(spec, w)
        p = prompts[b:b + 1].to(DEV)
        eng.reset_cache()
        first = eng.gen_next_token(p)
        toks = torch.cat([first.view(-1), eng.decode(n_new - 1)]).cpu()
        eng.reset_cache()
        logits = [eng.forward(p)[0, -1].float().cpu()]
        for i in range(n_new - 1):
            logits.append(eng.forward(toks[i].view(1, 1).to(DEV))[0, -1].float().cpu())
        out_logits.append(torch.stack(logits))
        out_toks.append(toks)
        eng.close()
    return torch.stack(out_logits, 1), torch.stack(out_toks, 1)          # [n, B, V], [n, B]


@pytest.mark.parametrize("spec", [models.TINY_QWEN2, models.TINY_LLAMA, models.TINY_QWEN3, models.TINY_MISTRAL],
                         ids=lambda s: s.name)
@pytest.mark.parametrize("B,S", [(2, 5), (4, 12), (8, 3), (3, 9)], ids=lambda v: str(v))
def test_batched_decode_is_bitwise_the_batch1_engine(built_lib, spec, B, S):
    w = {k: v.to(DEV) for k, v in models.synth_weights(spec, seed=5).items()}
    prompts = torch.randint(0, spec.vocab, (B, S), generator=torch.Generator().manual_seed(B * 100 + S))
    n_new = 10
    want_logits, want_toks = _single_runs(spec, w, prompts, n_new)
    eng = engine.DecodeEngine(spec, w)
    eng.reset_cache()
    first = eng.gen_next_token(prompts.to(DEV))                          # [B, 1]
    rest = eng.decode(n_new - 1)                                         # [n-1, B]
    toks = torch.cat([first.view(1, B), rest.view(n_new - 1, B)]).cpu()
    assert torch.equal(toks, want_toks), "batched greedy ids differ from the batch-1 engine's"
    # logits, teacher-forced on those tokens, step by step through forward([B, 1])
    eng.reset_cache()
    logits = [eng.forward(prompts.to(DEV))[:, -1].float().cpu()]
    for i in range(n_new - 1):
        logits.append(eng.forward(toks[i].view(B, 1).to(DEV))[:, -1].float().cpu())
    logits = torch.stack(logits)
    assert torch.equal(logits, want_logits), f"max |Δ| {float((logits - want_logits).abs().max())}"
    # all-position logits ([B, S, V]; token by token through the batched graph) = the batch-1 engine's for every sequence
    eng.reset_cache()
    allp = eng.forward(prompts.to(DEV), all_positions=True).float().cpu()
    assert allp.shape == (B, S, spec.vocab)
    one = engine.DecodeEngine(spec, w)
    for b in (0, B - 1):
        one.reset_cache()
        assert torch.equal(allp[b], one.forward(prompts[b:b + 1].to(DEV), all_positions=True)[0].float().cpu())
    one.close()
    # host-buffer API
    eng2 = eng.generate_sync_batch(prompts, n_new)
    assert torch.equal(eng2, want_toks.t())
    # ragged prompts are aligned the reference's way (left padding; the pads are attended, GPTEngine.cpp:95) — the same
    # batch as the padded ids given directly
    ragged = [prompts[b, min(b, S - 1):].tolist() for b in range(B)]
    padded, _ = engine.align_prompts(ragged, spec.max_ctx, 7)
    assert torch.equal(eng.generate_sync_batch(ragged, 4, pad_token=7), eng.generate_sync_batch(padded, 4))
    eng.close()


@pytest.mark.parametrize("B", [5, 8])
def test_batched_decode_wide_ffn_splits_the_down_projection(built_lib, B):
    """Mistral-7B's FFN width (k = 14336 for the down projection): 8 activation vectors of that length do not fit next
    to a useful ring, so the batched GEMV runs as launches of 4 sequences (csrc/gemv.cu gemv_plan_set_batch, `sub`) — same
    per-sequence arithmetic, so still bit-identical to batch-1 runs, uneven last group (B = 5) included."""
    spec = models.ModelSpec("tiny-wide-ffn", "mistral", 256, 2, 4, 2, 64, 14336, 512, 1e6, 1e-5, tie=False, max_ctx=64)
    w = {k: v.to(DEV) for k, v in models.synth_weights(spec, seed=9).items()}
    prompts = torch.randint(0, spec.vocab, (B, 6), generator=torch.Generator().manual_seed(B))
    n_new = 6
    want_logits, want_toks = _single_runs(spec, w, prompts, n_new)
    eng = engine.DecodeEngine(spec, w)
    eng.reset_cache()
    first = eng.gen_next_token(prompts.to(DEV))
    toks = torch.cat([first.view(1, B), eng.decode(n_new - 1).view(n_new - 1, B)]).cpu()
    assert torch.equal(toks, want_toks)
    eng.reset_cache()
    logits = [eng.forward(prompts.to(DEV))[:, -1].float().cpu()]
    for i in range(n_new - 1):
        logits.append(eng.forward(toks[i].view(B, 1).to(DEV))[:, -1].float().cpu())
    assert torch.equal(torch.stack(logits), want_logits)
    assert eng.launches_per_token > 1 + 5 * spec.layers + 2 - 1      # the extra down-projection launches are counted
    eng.close()


def test_batched_decode_full_size_and_weight_passes(built_lib):
    """Qwen2.5-0.5B at full size, B = 4 (the reference CLI's four prompts): bit-identical to four batch-1 runs, and a
    batched step costs far less than four batch-1 steps (the weights are streamed once)."""
    spec = models.QWEN25_05B.with_ctx(160)
    w = models.synth_weights(spec, seed=0, device=DEV, device_generator=True)
    B, S, n_new = 4, 16, 8
    prompts = torch.randint(0, spec.vocab, (B, S), generator=torch.Generator().manual_seed(7))
    want_logits, want_toks = _single_runs(spec, w, prompts, n_new)
    eng = engine.DecodeEngine(spec, w)
    eng.reset_cache()
    first = eng.gen_next_token(prompts.to(DEV))
    toks = torch.cat([first.view(1, B), eng.decode(n_new - 1).view(n_new - 1, B)]).cpu()
    assert torch.equal(toks, want_toks)
    # timing: 64 batched steps vs 64 batch-1 steps
    def timed(e, n=64):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e.decode(8)
        torch.cuda.synchronize()
        e0.record()
        e.decode(n)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    us_b = timed(eng)
    one = engine.DecodeEngine(spec, w)
    one.reset_cache()
    one.gen_next_token(prompts[:1].to(DEV))
    us_1 = timed(one)
    print(f"[Qwen2.5-0.5B] batched step (B = {B}): {us_b:.1f} us = {B / us_b * 1e6:.0f} tok/s; batch-1 step: {us_1:.1f} us = "
          f"{1 / us_1 * 1e6:.0f} tok/s; {B} sequential batch-1 engines would take {B * us_1:.1f} us")
    # measured: 997 µs for B = 4 against 4 × 411 µs — the small model's step is bound by instruction latency in the
    # consumer warps, and a batched stage issues B × the FMAs + x unpacks per weight vector
    assert us_b < 0.75 * B * us_1
    one.close()
    eng.close()


def test_batched_through_the_adapter_inside_the_reference(built_lib):
    """The reference program run on a [B, S] batch (its generateSync shape) plain and with our engine behind
    GPTModel::model(): the adapter's batched path equals our Python-driven batched engine bit for bit and stays inside
    the summation-order floor of the plain reference."""
    import sys
    import tempfile
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    import ref_cuda_parity as rp
    if not rp.REF_BIN.exists():
        pytest.skip("oracle/_ref/ref_cuda_decode not built (needs /root/reference: make -C oracle cuda)")
    spec = models.TINY_QWEN2
    w = models.synth_weights(spec, seed=0)
    B, S, n = 4, 16, 8
    prompts = torch.randint(0, spec.vocab, (B, S), generator=torch.Generator().manual_seed(3))
    eng = engine.DecodeEngine(spec, {k: v.to(DEV) for k, v in w.items()})
    eng.reset_cache()
    first = eng.gen_next_token(prompts.to(DEV))
    toks = torch.cat([first.view(1, B), eng.decode(n - 1).view(n - 1, B)]).cpu()
    eng.reset_cache()
    logits = [eng.forward(prompts.to(DEV))[:, -1].float().cpu()]
    for i in range(n - 1):
        logits.append(eng.forward(toks[i].view(B, 1).to(DEV))[:, -1].float().cpu())
    logits = torch.stack(logits)
    eng.close()
    with tempfile.TemporaryDirectory() as td:
        models.save_checkpoint(spec, w, td)
        flat = prompts.reshape(-1).tolist()
        forced = toks.reshape(-1).tolist()
        t_ref, l_ref, _ = rp.run_reference(spec, td, flat, n, forced=forced, batch=B)
        t_eng, l_eng, _ = rp.run_reference(spec, td, flat, n, forced=forced, batch=B, b200="engine")
    assert torch.equal(l_eng, logits) and torch.equal(t_eng, toks), "adapter (batched) != Python engine (batched)"
    d = (l_eng - l_ref).abs()
    ulp = 2.0 ** (torch.floor(torch.log2(l_ref.abs().max())).item() - 7)
    print(f"[tiny-qwen2 B={B}] reference program batched: ours vs plain mean {float(d.mean()):.3e} max {float(d.max()):.3e}")
    assert float(d.mean()) <= 4e-3 and float(d.max()) <= 8 * ulp
