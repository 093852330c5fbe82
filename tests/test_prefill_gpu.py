"""Batched prefill building blocks on the GPU: tensor-core causal attention (mma.sync, csrc/prefill_attn.cu) and the
persistent 128x256 tcgen05 GEMM (csrc/gemm.cu), each against the oracle and against the kernel it can replace
(B200_PREFILL_ATTN = mma | cuda, B200_GEMM = persistent | tile select either side whatever the default is)."""
import pytest
import torch

from helpers import assert_close_bf16, orc, to_oracle_cfg
from tinygpt_b200 import engine, models

pytestmark = pytest.mark.gpu
DEV = "cuda"

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


