"""Opt-in code paths that were written in a GPU-less session and have NOT yet run on hardware.

Skipped unless B200_STAGED=1 (tests/conftest.py), so the default `pytest -m gpu` run only contains verified paths.
Round 2 starts by running `B200_STAGED=1 python -m pytest tests/test_staged_gpu.py -m gpu -x -q` and promoting what
passes into the regular files (and, if faster, into the default path).

  * B200_FLAGSYNC=1 — per-op completion counters instead of griddepcontrol.wait between the kernels of a token
    (csrc/common.cuh FlagSync).  Same arithmetic, same kernels otherwise ⇒ logits and ids must be BIT-IDENTICAL to the
    default engine's, for body-only tokens (g_body graph), head tokens (g_step), rewinds and the batched prefill.
  * B200_PREFILL_ATTN=mma — tensor-core (mma.sync m16n8k16 bf16) causal flash attention for the batched prefill
    (csrc/prefill_attn.cu) against the oracle's causal attention with the reference's bf16-P rounding.
  * sampling kernels (csrc/sampling.cu) against the oracle's restatement of Sampler::sample.
"""
import pytest
import torch

from helpers import assert_close_bf16, orc, to_oracle_cfg
from tinygpt_b200 import engine, models

pytestmark = [pytest.mark.gpu, pytest.mark.staged]
DEV = "cuda"


def _run(spec, prompt, n_new, seed=0):
    """prompt token by token is not wanted here: keep the batched prefill out (S < 8) or in (S >= 8) via len(prompt)."""
    w = {k: v.to(DEV) for k, v in models.synth_weights(spec, seed=seed).items()}
    eng = engine.DecodeEngine(spec, w)
    out = {}
    eng.reset_cache()
    out["all_logits"] = eng.forward(prompt.view(1, -1).to(DEV), all_positions=True).float().cpu()  # g_step per token
    eng.reset_cache()
    first = eng.gen_next_token(prompt.view(1, -1).to(DEV))     # S-1 body-only tokens (g_body) + one head token
    rest = eng.decode(n_new - 1)
    out["toks"] = torch.cat([first.view(-1), rest]).cpu()
    eng.seek(len(prompt) + 3)                                  # rewind inside the generated suffix and regenerate
    again = eng.gen_next_token(out["toks"][3].view(1, 1).to(DEV))
    out["again"] = torch.cat([again.view(-1), eng.decode(4)]).cpu()
    out["launches"] = eng.launches_per_token
    torch.cuda.synchronize()
    eng.close()
    return out


@pytest.mark.parametrize("spec", [models.TINY_QWEN2, models.TINY_LLAMA, models.TINY_QWEN3, models.TINY_MISTRAL],
                         ids=lambda s: s.name)
@pytest.mark.parametrize("prompt_len", [5, 24], ids=["token-prefill", "gemm-prefill"])
def test_flagsync_engine_is_bit_identical(built_lib, spec, prompt_len, monkeypatch):
    prompt = torch.randint(0, spec.vocab, (prompt_len,), generator=torch.Generator().manual_seed(1))
    monkeypatch.setenv("B200_FLAGSYNC", "0")
    ref = _run(spec, prompt, 24)
    monkeypatch.setenv("B200_FLAGSYNC", "1")
    got = _run(spec, prompt, 24)
    assert got["launches"] == ref["launches"]
    assert torch.equal(got["all_logits"], ref["all_logits"]), "flag-sync changes no arithmetic: logits must be identical"
    assert torch.equal(got["toks"], ref["toks"])
    assert torch.equal(got["again"], ref["again"])
    assert torch.equal(got["again"], ref["toks"][4:9]), "rewind + regenerate reproduces the same suffix"


def test_flagsync_full_size_repeatable(built_lib, monkeypatch):
    """Qwen2.5-0.5B at full size: 64 tokens twice from the same state under B200_FLAGSYNC=1 → identical ids, and
    identical to the default engine (a missed dependency shows up as run-to-run differences)."""
    spec = models.QWEN25_05B.with_ctx(160)
    prompt = torch.randint(0, spec.vocab, (16,), generator=torch.Generator().manual_seed(0))
    runs = []
    for flag in ("0", "1", "1"):
        monkeypatch.setenv("B200_FLAGSYNC", flag)
        w = models.synth_weights(spec, seed=0, device=DEV, device_generator=True)
        eng = engine.DecodeEngine(spec, w)
        eng.reset_cache()
        first = eng.gen_next_token(prompt.view(1, -1).to(DEV))
        runs.append(torch.cat([first.view(-1), eng.decode(63)]).cpu())
        eng.close()
        del w
    assert torch.equal(runs[0], runs[1]) and torch.equal(runs[1], runs[2])


# ------------------------------------------------------------------------------ tensor-core prefill attention (mma.sync)
def _urand(*shape, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(*shape, generator=g) * 2 - 1).to(torch.bfloat16)


@pytest.mark.parametrize("B,S,Hq,Hkv,hd", [(2, 64, 4, 4, 128), (2, 67, 4, 2, 64), (1, 128, 14, 2, 64), (1, 1, 2, 1, 64),
                                            (3, 200, 8, 8, 128), (1, 513, 16, 8, 128), (1, 2, 1, 1, 128)])
def test_mma_causal_attention_vs_oracle(built_lib, B, S, Hq, Hkv, hd, monkeypatch):
    from tinygpt_b200 import ops
    q, k, v = _urand(B, S, Hq, hd, seed=42), _urand(B, S, Hkv, hd, seed=43), _urand(B, S, Hkv, hd, seed=44)
    monkeypatch.setenv("B200_PREFILL_ATTN", "mma")
    got = ops.flash_attention(q.to(DEV), k.to(DEV), v.to(DEV), True).float().cpu()
    monkeypatch.setenv("B200_PREFILL_ATTN", "cuda")
    base = ops.flash_attention(q.to(DEV), k.to(DEV), v.to(DEV), True).float().cpu()   # verified CUDA-core kernel
    naive = orc.naive_attention(q, k, v, True)
    err = (got - naive).abs()
    assert bool((err <= 1e-2 + 1e-1 * naive.abs()).all()), f"TinyFA bf16 tolerance violated: max {float(err.max())}"
    want = orc.flash_attention(q, k, v, True)
    assert_close_bf16(got, want, 4, "mma prefill attention vs oracle tile walk", atol=4e-3)
    assert_close_bf16(got, base, 6, "mma prefill attention vs the CUDA-core kernel", atol=6e-3)


@pytest.mark.parametrize("spec,S", [(models.TINY_QWEN2, 9), (models.TINY_QWEN3, 300), (models.TINY_LLAMA, 600),
                                    (models.TINY_MISTRAL, 129)], ids=lambda v: getattr(v, "name", str(v)))
def test_mma_prefill_through_engine(built_lib, spec, S, monkeypatch):
    """Whole batched prefill with the tensor-core attention vs the oracle's causal prefill, then a second chunk at an
    offset (mask shifted by the cached prefix) vs one-shot."""
    spec = spec.with_ctx(1024)
    w = {k: v.to(DEV) for k, v in models.synth_weights(spec, seed=4).items()}
    prompt = torch.randint(0, spec.vocab, (1, S), generator=torch.Generator().manual_seed(S)).to(DEV)
    monkeypatch.setenv("B200_PREFILL_ATTN", "mma")
    eng = engine.DecodeEngine(spec, w)
    eng.reset_cache()
    got = eng.forward(prompt)[0, -1].float().cpu()
    toks = eng.decode(8).cpu()
    eng.reset_cache()
    cut = max(8, S // 3)
    if S - cut >= 8:
        eng.forward(prompt[:, :cut])
        two = eng.forward(prompt[:, cut:])[0, -1].float().cpu()
        assert float((two - got).abs().max()) <= 2e-2 and float((two - got).abs().mean()) <= 2e-3
    monkeypatch.setenv("B200_PREFILL_ATTN", "cuda")
    eng.reset_cache()
    base = eng.forward(prompt)[0, -1].float().cpu()
    base_toks = eng.decode(8).cpu()
    cfg = to_oracle_cfg(spec)
    wc = {k: v.float().cpu() for k, v in w.items()}
    want = orc.forward(cfg, wc, prompt.cpu(), orc.KVCache(), models.rope_table(spec), "bf16")[0, -1]
    top = float(want.abs().max())
    ulp = 2.0 ** (torch.floor(torch.log2(torch.tensor(top))).item() - 7)
    d = (got - want).abs()
    print(f"[{spec.name} S={S}] mma prefill vs oracle: mean {float(d.mean()):.3e} max {float(d.max()):.3e}; "
          f"vs CUDA-core prefill max {float((got - base).abs().max()):.3e}; ids equal {int((toks == base_toks).sum())}/8")
    assert float(d.mean()) <= 4e-3 and float(d.max()) <= 8 * ulp
    eng.close()


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


# ---------------------------------------------------------------------- persistent 128×256 tcgen05 GEMM (B200_GEMM=persistent)
def _rand_bf16(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 128, 256), (256, 384, 512), (16, 1152, 896), (200, 896, 4864),
                                   (1, 128, 64), (130, 136, 72), (2048, 4096, 2048), (512, 12288, 2048),
                                   (2048, 2048, 6144), (384, 40000, 128)])
def test_persistent_gemm_vs_oracle_and_v1(built_lib, M, N, K, monkeypatch):
    """Same gate as tests/test_gemm_gpu.py, and equality with the verified one-tile-per-CTA kernel: both accumulate the
    same bf16 products in fp32 on the same tensor cores in the same k order, so they should agree bit for bit."""
    from tinygpt_b200 import ops
    a, w = _rand_bf16(M, K, seed=M + K), _rand_bf16(N, K, seed=N, scale=0.02)
    monkeypatch.setenv("B200_GEMM", "persistent")
    got = ops.gemm(a.to(DEV), w.to(DEV))
    got2 = ops.gemm(a.to(DEV), w.to(DEV))            # TMEM buffers / barriers re-used correctly on a second launch
    monkeypatch.setenv("B200_GEMM", "tile")
    base = ops.gemm(a.to(DEV), w.to(DEV))
    assert torch.equal(got, got2)
    if M * N * K <= 2 ** 31:
        want = orc.linear(a, w, None)
    else:
        want = (a.to(DEV).float() @ w.to(DEV).float().t()).to(torch.bfloat16).float().cpu()
    assert_close_bf16(got, want, 1, f"persistent gemm {M}x{N}x{K}", atol=1e-4, frac_exact=0.97)
    same = float((got == base).float().mean())
    print(f"[persistent gemm {M}x{N}x{K}] bit-identical to the 128x128 kernel on {same:.5f} of the elements")
    assert same >= 0.999


def test_prefill_chunk_override(built_lib, monkeypatch):
    """B200_PREFILL_CHUNK=2048: one pass instead of chunks of 512 — same logits up to summation order."""
    spec = models.TINY_LLAMA.with_ctx(2304)
    w = {k: v.to(DEV) for k, v in models.synth_weights(spec, seed=4).items()}
    prompt = torch.randint(0, spec.vocab, (1, 1500), generator=torch.Generator().manual_seed(5)).to(DEV)
    eng = engine.DecodeEngine(spec, w)
    eng.reset_cache()
    a = eng.forward(prompt)[0, -1].float().cpu()
    monkeypatch.setenv("B200_PREFILL_CHUNK", "2048")
    eng.reset_cache()
    b = eng.forward(prompt)[0, -1].float().cpu()
    assert float((a - b).abs().max()) <= 2e-2 and float((a - b).abs().mean()) <= 2e-3
    eng.close()


# --------------------------------------------------------------------------------- loader fast path, CUDA destination
def test_loader_cuda_staging_path(built_lib, tmp_path):
    """The pinned-staging H2D path (chunked, strided column slices included) delivers the same bytes as the CPU path,
    and the loaded tensors drive the engine to the same tokens as the in-memory synthetic checkpoint."""
    from tinygpt_b200 import loader, tp
    spec = models.TINY_QWEN2
    w = models.synth_weights(spec, seed=9)
    models.save_checkpoint(spec, w, str(tmp_path))
    got_spec, got, _ = loader.load_checkpoint(tmp_path, device=DEV)
    for k, v in w.items():
        assert torch.equal(got[k].cpu().view(torch.int16), v.view(torch.int16)), k
    _, r1, _ = loader.load_checkpoint(tmp_path, device=DEV, rank=1, world=2)
    want = tp.shard_weights(spec, w, 1, 2)
    for k, v in want.items():
        assert torch.equal(r1[k].cpu().view(torch.int16), v.view(torch.int16)), k
    prompt = [3, 1, 4, 1, 5, 9, 2, 6, 5]
    a = engine.DecodeEngine(got_spec.with_ctx(128), got)
    b = engine.DecodeEngine(spec.with_ctx(128), {k: v.to(DEV) for k, v in w.items()})
    assert a.generate_sync(prompt, 16).tolist() == b.generate_sync(prompt, 16).tolist()
    a.close()
    b.close()


# ------------------------------------------------------------------------------- cross-kernel L2 prefetch (B200_L2PF_MB)
@pytest.mark.parametrize("flags", [{"B200_L2PF_MB": "16"}, {"B200_L2PF_MB": "24", "B200_FLAGSYNC": "1"}],
                         ids=["l2pf", "l2pf+flagsync"])
def test_l2_prefetch_changes_nothing_but_time(built_lib, flags, monkeypatch):
    """The prefetch is a hint: ids must be identical with it on (full-size Qwen2.5-0.5B: the lm_head's 272 MB and the
    per-layer matrices beyond the rings are prefetched by the preceding kernels; a bad address range would fault)."""
    spec = models.QWEN25_05B.with_ctx(160)
    prompt = torch.randint(0, spec.vocab, (16,), generator=torch.Generator().manual_seed(0))
    runs = []
    for on in (False, True):
        for k, v in flags.items():
            monkeypatch.setenv(k, v if on else "0")
        w = models.synth_weights(spec, seed=0, device=DEV, device_generator=True)
        eng = engine.DecodeEngine(spec, w)
        eng.reset_cache()
        first = eng.gen_next_token(prompt.view(1, -1).to(DEV))
        runs.append(torch.cat([first.view(-1), eng.decode(47)]).cpu())
        torch.cuda.synchronize()
        eng.close()
        del w
    assert torch.equal(runs[0], runs[1])


# --------------------------------------------------------------------- flag-sync + L2 prefetch on tensor-parallel engines
def _tp_fs_worker(rank, world, port, spec_name, shard_attn, q):
    import os
    import torch.distributed as dist
    from tinygpt_b200 import tp
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        spec = models.SPECS[spec_name]
        if spec.max_ctx > 512:
            spec = spec.with_ctx(256)
        w = models.synth_weights(spec, seed=0)
        prompt = torch.randint(0, spec.vocab, (1, 9), generator=torch.Generator().manual_seed(0)).to(dev)
        out = []
        for flags in ({"B200_FLAGSYNC": "0", "B200_FLAGSYNC_TP": "0", "B200_L2PF_MB": "0"},
                      {"B200_FLAGSYNC": "1", "B200_FLAGSYNC_TP": "1", "B200_L2PF_MB": "8"}):
            os.environ.update(flags)
            eng = tp.TPDecodeEngine(spec, w, rank, world, dev, shard_attn=shard_attn)
            eng.reset_cache()
            first = eng.gen_next_token(prompt)              # 8 body-only tokens + 1 head token, then 23 head tokens
            toks = torch.cat([first.view(-1), eng.decode(23)]).cpu()
            eng.reset_cache()
            local = eng.forward(prompt)[0, -1].float().cpu()
            out.append((toks, local))
            dist.barrier()
            eng.close()
        if rank == 0:
            q.put(out)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("spec_name,shard_attn", [("tiny-mistral", True), ("tiny-qwen2", False), ("Qwen2.5-0.5B", True)])
def test_tp2_flagsync_bit_identical(built_lib, spec_name, shard_attn):
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_tp_fs_worker, args=(r, 2, port, spec_name, shard_attn, q)) for r in range(2)]
    for p in procs:
        p.start()
    (toks_a, logits_a), (toks_b, logits_b) = q.get(timeout=200)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert torch.equal(toks_a, toks_b) and torch.equal(logits_a, logits_b)


# ------------------------------------------------------------------ register-resident small-k GEMV loop (B200_GEMV_SMALLK)
@pytest.mark.parametrize("n,k,kind", [(1152, 896, "norm_bias"), (896, 896, "residual"), (4864, 896, "silu"),
                                      (1152, 896, "norm"), (256, 896, "silu"), (40, 264, "residual"), (4000, 1024, "silu"),
                                      (8, 8, "norm"), (5000, 512, "silu")])
def test_smallk_gemv_is_bit_identical(built_lib, n, k, kind, monkeypatch):
    """Same FMAs in the same order per row ⇒ the outputs must equal the default loop's bit for bit."""
    from tinygpt_b200 import ops
    g = torch.Generator().manual_seed(n + k)
    x = torch.randn(k, generator=g).to(torch.bfloat16).to(DEV)
    nw = (1 + 0.02 * torch.randn(k, generator=g)).to(torch.bfloat16).to(DEV)
    rows = 2 * n if kind == "silu" else n
    w = (0.02 * torch.randn(rows, k, generator=g)).to(torch.bfloat16).to(DEV)
    vec = torch.randn(n, generator=g).to(torch.bfloat16).to(DEV)

    def run():
        if kind == "norm_bias":
            return ops.gemv_fused(x, w, norm_weight=nw, eps=1e-6, bias=vec)
        if kind == "norm":
            return ops.gemv_fused(x, w, norm_weight=nw, eps=1e-6)
        if kind == "residual":
            return ops.gemv_fused(x, w, residual=vec)
        return ops.gemv_fused(x, w, norm_weight=nw, eps=1e-6, silu_mul=True)

    monkeypatch.setenv("B200_GEMV_SMALLK", "0")
    base = run()
    monkeypatch.setenv("B200_GEMV_SMALLK", "1")
    got = run()
    torch.cuda.synchronize()
    assert torch.equal(got, base)


@pytest.mark.parametrize("flags", [{"B200_GEMV_SMALLK": "1"}, {"B200_GEMV_SMALLK": "1", "B200_FLAGSYNC": "1"}],
                         ids=["smallk", "smallk+flagsync"])
def test_smallk_engine_bit_identical(built_lib, flags, monkeypatch):
    for spec in (models.TINY_QWEN2, models.QWEN25_05B.with_ctx(160)):
        prompt = torch.randint(0, spec.vocab, (6,), generator=torch.Generator().manual_seed(2))
        runs = []
        for on in (False, True):
            for k, v in flags.items():
                monkeypatch.setenv(k, v if on else "0")
            w = models.synth_weights(spec, seed=0, device=DEV, device_generator=True)
            eng = engine.DecodeEngine(spec, w)
            eng.reset_cache()
            logits = eng.forward(prompt.view(1, -1).to(DEV), all_positions=True).float().cpu()
            toks = eng.decode(20).cpu()
            torch.cuda.synchronize()
            eng.close()
            del w
            runs.append((logits, toks))
        assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1]), spec.name


# ------------------------------------------------------------- the reference's own CUDA path as the oracle (oracle/_ref)
def test_engine_and_oracle_against_reference_cuda(built_lib):
    """oracle/_ref/ref_cuda_decode = the UNMODIFIED reference CUDA build (TinyTorch ops + cuBLAS + TinyFA), compiled in
    the container that has /root/reference (`make -C oracle cuda`).  Same synthetic checkpoint, same forced tokens:
    engine vs reference and oracle vs reference within the summation-order floor (cuBLAS' order is not ours), greedy
    ids equal wherever the reference's own top-2 margin is decisive."""
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    sys.path.insert(0, str(root / "tools"))
    import ref_cuda_parity as rp
    if not rp.REF_BIN.exists():
        pytest.skip("oracle/_ref/ref_cuda_decode not built (needs /root/reference: make -C oracle cuda)")
    import tempfile
    for spec in (models.TINY_QWEN2, models.TINY_LLAMA, models.TINY_QWEN3, models.TINY_MISTRAL):
        w = models.synth_weights(spec, seed=0)
        prompt = torch.randint(0, spec.vocab, (16,), generator=torch.Generator().manual_seed(0)).tolist()
        with tempfile.TemporaryDirectory() as td:
            models.save_checkpoint(spec, w, td)
            toks, logits, _ = rp.run_engine(spec, w, prompt, 16)
            ref_toks, ref_logits, _ = rp.run_reference(spec, td, prompt, 16, forced=toks.tolist())
        wf = {k: v.float() for k, v in w.items()}
        _, logits_orc = orc.generate_greedy(to_oracle_cfg(spec), wf, torch.tensor(prompt), 16, models.rope_table(spec),
                                            "bf16", forced=toks)
        d_eng, d_orc = (logits - ref_logits).abs(), (logits_orc - ref_logits).abs()
        top = float(ref_logits.abs().max())
        ulp = 2.0 ** (torch.floor(torch.log2(torch.tensor(top))).item() - 7)
        print(f"[{spec.name}] engine-vs-reference-CUDA mean {float(d_eng.mean()):.3e} max {float(d_eng.max()):.3e}; "
              f"oracle-vs-reference-CUDA mean {float(d_orc.mean()):.3e} max {float(d_orc.max()):.3e}; ulp {ulp:.3e}")
        assert float(d_orc.mean()) <= 4e-3 and float(d_orc.max()) <= 8 * ulp, "oracle's bf16 rounding points are off"
        assert float(d_eng.mean()) <= 4e-3 and float(d_eng.max()) <= 8 * ulp
        srt = torch.sort(ref_logits, dim=-1, descending=True).values
        decided = (srt[:, 0] - srt[:, 1]) > 4 * ulp
        assert torch.equal(ref_toks[decided], toks[decided]), f"{spec.name}: greedy ids differ on decisive steps"


def test_drop_in_boundary_inside_the_real_reference(built_lib):
    """The SAME reference program (its loader, modules, KV manager, generate loop, argmax) with
    (a) our engine behind GPTModel::model() via b200::adapter::ModelB200 — must reproduce our Python-driven engine bit
        for bit (same library, same weights, same prefill path), and
    (b) our kernels behind its op registry via b200::adapter::registerOps() — must stay within the summation-order floor
        of the plain reference."""
    import sys
    import tempfile
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    import ref_cuda_parity as rp
    if not rp.REF_BIN.exists():
        pytest.skip("oracle/_ref/ref_cuda_decode not built (needs /root/reference: make -C oracle cuda)")
    for spec in (models.TINY_QWEN2, models.TINY_QWEN3, models.TINY_MISTRAL):
        w = models.synth_weights(spec, seed=0)
        prompt = torch.randint(0, spec.vocab, (16,), generator=torch.Generator().manual_seed(0)).tolist()
        with tempfile.TemporaryDirectory() as td:
            models.save_checkpoint(spec, w, td)
            toks, logits, _ = rp.run_engine(spec, w, prompt, 12)
            _, ref_logits, _ = rp.run_reference(spec, td, prompt, 12, forced=toks.tolist())
            t_eng, l_eng, _ = rp.run_reference(spec, td, prompt, 12, forced=toks.tolist(), b200="engine")
            t_ops, l_ops, _ = rp.run_reference(spec, td, prompt, 12, forced=toks.tolist(), b200="ops")
        assert torch.equal(l_eng, logits) and torch.equal(t_eng, toks), f"{spec.name}: adapter engine != Python engine"
        top = float(ref_logits.abs().max())
        ulp = 2.0 ** (torch.floor(torch.log2(torch.tensor(top))).item() - 7)
        d = (l_ops - ref_logits).abs()
        print(f"[{spec.name}] reference + our ops vs plain reference: mean {float(d.mean()):.3e} max {float(d.max()):.3e}")
        assert float(d.mean()) <= 4e-3 and float(d.max()) <= 8 * ulp


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
