"""Attention parity: (a) the general kernel behind b200_attn_bf16 on the shapes of TinyFA's own test-suite
(TFAroot/tests/cpp/test_flash_attn.cu:350-403, head dims 64/128 — the only ones TinyGPT compiles,
src/CMakeLists.txt:17-23) with TinyFA's bf16 tolerance (rel 1e-1 / abs 1e-2, :92-110) against the naive fp32
attention of its cpu_reference.h AND, tighter, against the oracle's tile-walk restatement; (b) the fused split-KV
decode kernel the engine runs (q/k-norm + RoPE + in-place append + attention + merge)."""
import pytest
import torch

from helpers import assert_close_bf16, orc
from tinygpt_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def urand_bf16(*shape, seed=42):
    g = torch.Generator().manual_seed(seed)  # TinyFA tests: U(-1, 1), seed 42
    return (torch.rand(*shape, generator=g) * 2 - 1).to(torch.bfloat16)


TFA_SHAPES = [  # (B, Sq, Skv, Hq, Hkv, hd, causal)
    (2, 64, 64, 4, 4, 64, False), (2, 256, 256, 8, 8, 128, False), (2, 64, 64, 4, 4, 128, False),
    (2, 64, 64, 4, 4, 128, True), (2, 1, 16, 4, 4, 64, False), (2, 32, 128, 4, 4, 64, False),
    (2, 128, 32, 4, 4, 64, False), (2, 67, 83, 4, 4, 64, False), (3, 97, 101, 4, 4, 64, False),
    (1, 512, 512, 8, 8, 64, False), (8, 64, 64, 4, 4, 64, False), (2, 64, 64, 1, 1, 64, False),
    (2, 64, 64, 16, 16, 64, False), (2, 64, 64, 8, 4, 64, False), (2, 64, 64, 8, 2, 64, False),
    (2, 64, 64, 8, 1, 64, False), (2, 67, 83, 4, 2, 64, True), (1, 128, 128, 14, 2, 64, True),
]


@pytest.mark.parametrize("B,Sq,Skv,Hq,Hkv,hd,causal", TFA_SHAPES)
def test_attention_tinyfa_shapes(built_lib, B, Sq, Skv, Hq, Hkv, hd, causal):
    q, k, v = urand_bf16(B, Sq, Hq, hd, seed=42), urand_bf16(B, Skv, Hkv, hd, seed=43), urand_bf16(B, Skv, Hkv, hd, seed=44)
    got = ops.flash_attention(q.to(DEV), k.to(DEV), v.to(DEV), causal).float().cpu()
    naive = orc.naive_attention(q, k, v, causal)
    err = (got - naive).abs()
    assert bool((err <= 1e-2 + 1e-1 * naive.abs()).all()), f"TinyFA bf16 tolerance violated: max {float(err.max())}"
    # against the restated reference kernel (bf16 P rounding modelled): a few bf16 ulps of the output
    want = orc.flash_attention(q, k, v, causal)
    assert_close_bf16(got, want, 4, "attention vs oracle tile walk", atol=4e-3)


@pytest.mark.parametrize("Hq,Hkv,hd,L", [(14, 2, 64, 144), (24, 8, 128, 240), (16, 8, 128, 31), (32, 8, 128, 129),
                                          (4, 4, 64, 1), (8, 1, 64, 500), (6, 2, 128, 17), (14, 2, 64, 257),
                                          (16, 8, 128, 512), (14, 2, 64, 256)])
@pytest.mark.parametrize("max_ctx", [512, 2048], ids=["cta-per-head", "cta-per-gqa-group"])
def test_attn_decode_plain(built_lib, Hq, Hkv, hd, L, max_ctx):
    """pos == NULL mode: attention of one query row over `L` cached rows (no rotation, nothing appended).
    max_ctx ≤ 1024 runs one CTA per query head, longer contexts one CTA per GQA group (K/V read once)."""
    q = urand_bf16(1, 1, Hq, hd, seed=1)
    k, v = urand_bf16(1, L, Hkv, hd, seed=2), urand_bf16(1, L, Hkv, hd, seed=3)
    kc = torch.zeros(max_ctx, Hkv, hd, dtype=torch.bfloat16)
    vc = torch.zeros_like(kc)
    kc[:L], vc[:L] = k[0], v[0]
    qkv = torch.cat([q.view(-1), torch.zeros(2 * Hkv * hd, dtype=torch.bfloat16)])
    got = ops.attn_decode(qkv.to(DEV), kc.to(DEV), vc.to(DEV), q_heads=Hq, kv_heads=Hkv, head_dim=hd, fixed_len=L).float().cpu().view(1, 1, Hq, hd)
    want = orc.flash_attention(q, k, v, False)
    assert_close_bf16(got, want, 4, "split-KV decode attention", atol=4e-3)
    naive = orc.naive_attention(q, k, v, False)
    assert bool(((got - naive).abs() <= 1e-2 + 1e-1 * naive.abs()).all())


@pytest.mark.parametrize("max_ctx", [384, 1536], ids=["cta-per-head", "cta-per-gqa-group"])
@pytest.mark.parametrize("Hq,Hkv,hd,qk_norm", [(14, 2, 64, False), (24, 8, 128, False), (16, 8, 128, True),
                                               (32, 8, 128, False)])
def test_attn_decode_fused_append(built_lib, Hq, Hkv, hd, qk_norm, max_ctx):
    """The engine's launch: raw qkv → [norm] → RoPE → append at *pos → attention, for several consecutive positions.
    K/V rows written to the cache must be bit-identical to the oracle's rope(k) / v (same table, ≤ 1 ulp FMA)."""
    steps, start = 5, 254  # crosses the 256-key (hd 64) / 128-key (hd 128) split boundaries
    table = orc.rope_table(hd, max_ctx, 1e6)
    g = torch.Generator().manual_seed(7)
    kc = torch.zeros(max_ctx, Hkv, hd, dtype=torch.bfloat16)
    vc = torch.zeros_like(kc)
    kc[:start] = (torch.rand(start, Hkv, hd, generator=g) * 2 - 1).to(torch.bfloat16)
    vc[:start] = (torch.rand(start, Hkv, hd, generator=g) * 2 - 1).to(torch.bfloat16)
    qn = (1 + 0.1 * torch.randn(hd, generator=g)).to(torch.bfloat16) if qk_norm else None
    kn = (1 + 0.1 * torch.randn(hd, generator=g)).to(torch.bfloat16) if qk_norm else None
    kc_d, vc_d = kc.to(DEV), vc.to(DEV)
    Kref, Vref = kc[:start].clone().unsqueeze(0), vc[:start].clone().unsqueeze(0)
    pos = torch.zeros(1, dtype=torch.int32, device=DEV)
    for t in range(steps):
        p = start + t
        qkv = (torch.rand((Hq + 2 * Hkv) * hd, generator=g) * 2 - 1).to(torch.bfloat16)
        pos.fill_(p)
        got = ops.attn_decode(qkv.to(DEV), kc_d, vc_d, q_heads=Hq, kv_heads=Hkv, head_dim=hd, pos=pos,
                              rope_table=table.to(DEV), q_norm=None if qn is None else qn.to(DEV),
                              k_norm=None if kn is None else kn.to(DEV), eps=1e-6).float().cpu()
        q = qkv[: Hq * hd].view(1, 1, Hq, hd)
        k = qkv[Hq * hd: (Hq + Hkv) * hd].view(1, 1, Hkv, hd)
        v = qkv[(Hq + Hkv) * hd:].view(1, 1, Hkv, hd)
        if qk_norm:
            q, k = orc.rms_norm(q, qn, 1e-6), orc.rms_norm(k, kn, 1e-6)
        q, k = orc.rope_apply(q, table, p), orc.rope_apply(k, table, p)
        Kref, Vref = torch.cat([Kref, k], 1), torch.cat([Vref, v.float()], 1)
        want = orc.flash_attention(q, Kref, Vref, False)
        assert_close_bf16(got.view(1, 1, Hq, hd), want, 4, f"fused decode attention step {t}", atol=4e-3)
        assert_close_bf16(kc_d[p].float().cpu(), k[0, 0], 1, "appended K row", frac_exact=0.95)
        assert torch.equal(vc_d[p].float().cpu(), v[0, 0].float()), "appended V row must be an exact copy"
        Kref[0, -1] = kc_d[p].float().cpu()  # continue from the device's K so that a 1-ulp K difference cannot pile up
    assert torch.equal(kc_d[:start].cpu(), kc[:start]), "earlier cache rows must be untouched"
