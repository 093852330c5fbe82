"""Batched decode (B ≤ 8 sequences at the same position — GPTEngine::generateSync's left-padded batch,
src/engine/GPTEngine.cpp:101-174): one weight pass per step for the whole batch (csrc/gemv_batch.cu: the B activation vectors
are the n = 8 operand of tensor-core MMAs), attention with grid.z = B, per-sequence KV caches.  Everything but the
GEMV's k-sum is the batch-1 arithmetic, so a batched run must agree with B independent batch-1 runs inside the
summation-order floor — the gate of the TP-vs-single-GPU and engine-vs-reference-CUDA tests — with greedy ids equal
wherever the top-2 margin is decisive."""
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


def _close(got, want, what):
    """Tensor-core batched GEMV vs batch-1 runs: same gate as TP-vs-single-GPU and engine-vs-reference-CUDA — inside the
    summation-order floor (mean ≤ 0.5 ulp, max ≤ 8 ulp of the top logit for the tiny shapes)."""
    d = (got - want).abs()
    ulp = 2.0 ** (torch.floor(torch.log2(want.abs().max())).item() - 7)
    assert float(d.mean()) <= 0.5 * ulp and float(d.max()) <= 8 * ulp, \
        f"{what}: mean {float(d.mean()):.3e} max {float(d.max()):.3e} (1 ulp = {ulp:.3e})"
    return float(d.max())


def _decisive(want_logits, dmax):
    top2 = torch.topk(want_logits, 2, dim=-1).values
    return (top2[..., 0] - top2[..., 1]) > 2 * dmax


@pytest.mark.parametrize("spec", [models.TINY_QWEN2, models.TINY_LLAMA, models.TINY_QWEN3, models.TINY_MISTRAL],
                         ids=lambda s: s.name)
@pytest.mark.parametrize("B,S", [(2, 5), (4, 12), (8, 3), (3, 9)], ids=lambda v: str(v))
def test_batched_decode_matches_batch1(built_lib, spec, B, S):
    """Every sequence of a batched run against its own
    batch-1 run — logits of the prompt and of teacher-forced steps inside the summation-order floor, greedy ids equal
    wherever the batch-1 top-2 margin is decisive."""
    w = {k: v.to(DEV) for k, v in models.synth_weights(spec, seed=5).items()}
    prompts = torch.randint(0, spec.vocab, (B, S), generator=torch.Generator().manual_seed(B * 100 + S))
    n_new = 10
    want_logits, want_toks = _single_runs(spec, w, prompts, n_new)
    eng = engine.DecodeEngine(spec, w)
    eng.reset_cache()
    logits = [eng.forward(prompts.to(DEV))[:, -1].float().cpu()]
    for i in range(n_new - 1):
        logits.append(eng.forward(want_toks[i].view(B, 1).to(DEV))[:, -1].float().cpu())
    logits = torch.stack(logits)
    dmax = _close(logits, want_logits, f"{spec.name} B={B}")
    eng.reset_cache()
    first = eng.gen_next_token(prompts.to(DEV))
    toks = torch.cat([first.view(1, B), eng.decode(n_new - 1).view(n_new - 1, B)]).cpu()
    # free-running ids: equal up to the first step whose margin is not decisive (after it the sequences may part)
    dec = _decisive(want_logits, dmax)
    for b in range(B):
        upto = n_new if bool(dec[:, b].all()) else int((~dec[:, b]).nonzero()[0])
        assert torch.equal(toks[:upto, b], want_toks[:upto, b]), f"sequence {b}: ids differ on a decisive step"
    assert torch.equal(logits.argmax(-1)[dec], want_logits.argmax(-1)[dec])
    # deterministic, and the host-buffer API runs the same graph
    assert torch.equal(eng.generate_sync_batch(prompts, n_new), toks.t())
    # all-position logits ([B, S, V]; token by token through the batched graph) against the batch-1 engine's
    eng.reset_cache()
    allp = eng.forward(prompts.to(DEV), all_positions=True).float().cpu()
    assert allp.shape == (B, S, spec.vocab)
    one = engine.DecodeEngine(spec, w)
    for b in (0, B - 1):
        one.reset_cache()
        _close(allp[b], one.forward(prompts[b:b + 1].to(DEV), all_positions=True)[0].float().cpu(), "all positions")
    one.close()
    # ragged prompts are aligned the reference's way (left padding; the pads are attended, GPTEngine.cpp:95) — the same
    # batch as the padded ids given directly
    ragged = [prompts[b, min(b, S - 1):].tolist() for b in range(B)]
    padded, _ = engine.align_prompts(ragged, spec.max_ctx, 7)
    assert torch.equal(eng.generate_sync_batch(ragged, 4, pad_token=7), eng.generate_sync_batch(padded, 4))
    eng.close()


@pytest.mark.parametrize("B", [5, 8])
def test_batched_decode_wide_ffn_splits_the_down_projection(built_lib, B):
    """Mistral-7B's FFN width (k = 14336 for the down projection): 8 activation vectors of that length do not fit next
    to a useful ring, so the batched GEMV runs as launches of ≤ 4 sequences (csrc/gemv.cu gemv_plan_set_batch, `sub`),
    uneven last group (B = 5 → 3 + 2) included."""
    spec = models.ModelSpec("tiny-wide-ffn", "mistral", 256, 2, 4, 2, 64, 14336, 512, 1e6, 1e-5, tie=False, max_ctx=64)
    w = {k: v.to(DEV) for k, v in models.synth_weights(spec, seed=9).items()}
    prompts = torch.randint(0, spec.vocab, (B, 6), generator=torch.Generator().manual_seed(B))
    n_new = 6
    want_logits, want_toks = _single_runs(spec, w, prompts, n_new)
    eng = engine.DecodeEngine(spec, w)
    eng.reset_cache()
    first = eng.gen_next_token(prompts.to(DEV))
    toks = torch.cat([first.view(1, B), eng.decode(n_new - 1).view(n_new - 1, B)]).cpu()
    eng.reset_cache()
    logits = [eng.forward(prompts.to(DEV))[:, -1].float().cpu()]
    for i in range(n_new - 1):
        logits.append(eng.forward(want_toks[i].view(B, 1).to(DEV))[:, -1].float().cpu())
    dmax = _close(torch.stack(logits), want_logits, f"wide FFN B={B}")
    dec = _decisive(want_logits, dmax)
    assert torch.equal(toks[0][dec[0]], want_toks[0][dec[0]])
    assert eng.launches_per_token > 1 + 5 * spec.layers + 2 - 1      # the extra down-projection launches are counted
    eng.close()


def test_batched_decode_full_size_and_weight_passes(built_lib):
    """Qwen2.5-0.5B at full size, B = 4 (the reference CLI's four prompts) against four batch-1 runs — inside the
    full-size summation-order floor (≤ 1 ulp mean, ≤ 16 ulp max: the TP gate) — and a batched step costs far less than four batch-1 steps (the weights are streamed once)."""
    spec = models.QWEN25_05B.with_ctx(160)
    w = models.synth_weights(spec, seed=0, device=DEV, device_generator=True)
    B, S, n_new = 4, 16, 8
    prompts = torch.randint(0, spec.vocab, (B, S), generator=torch.Generator().manual_seed(7))
    want_logits, want_toks = _single_runs(spec, w, prompts, n_new)
    eng = engine.DecodeEngine(spec, w)
    eng.reset_cache()
    logits = [eng.forward(prompts.to(DEV))[:, -1].float().cpu()]
    for i in range(n_new - 1):
        logits.append(eng.forward(want_toks[i].view(B, 1).to(DEV))[:, -1].float().cpu())
    logits = torch.stack(logits)
    eng.reset_cache()
    first = eng.gen_next_token(prompts.to(DEV))
    toks = torch.cat([first.view(1, B), eng.decode(n_new - 1).view(n_new - 1, B)]).cpu()
    d = (logits - want_logits).abs()
    ulp = 2.0 ** (torch.floor(torch.log2(want_logits.abs().max())).item() - 7)
    print(f"[Qwen2.5-0.5B B={B}] logits vs batch-1 runs: mean {float(d.mean()):.3e} max {float(d.max()):.3e} "
          f"(1 ulp = {ulp:.3e}); ids equal on {int((toks == want_toks).sum())}/{toks.numel()}")
    assert float(d.mean()) <= 1.0 * ulp and float(d.max()) <= 16 * ulp
    dec = _decisive(want_logits, float(d.max()))
    assert torch.equal(logits.argmax(-1)[dec], want_logits.argmax(-1)[dec])
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
    # measured: ~1.4 × a batch-1 step at B = 4 (the per-kernel dependency latency of the small model does not shrink)
    assert us_b < 0.5 * B * us_1
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
