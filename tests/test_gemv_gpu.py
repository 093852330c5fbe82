"""Parity of the weight-streaming GEMV (plain and fused variants) against the oracle at every (n, k) of the decode
path plus ragged shapes.  Tolerance: ≤ 1 bf16 ulp (fp32 summation order differs from the oracle's
and from cuBLAS'; products are exact), or 1e-4 absolute near zero; ≥ 98 % of elements bit-identical."""
import pytest
import torch

from helpers import assert_close_bf16, orc
from tinygpt_b200 import models, ops
from tinygpt_b200._lib import B200Error

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rand_bf16(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16)


def shapes():
    out = []
    for s in (models.QWEN25_05B, models.LLAMA32_3B, models.QWEN3_17B, models.MISTRAL_7B):
        out += [(s.q_dim + 2 * s.kv_dim, s.hidden), (s.hidden, s.q_dim), (s.hidden, s.intermediate)]
    return sorted(set(out))


@pytest.mark.parametrize("n,k", shapes() + [(8, 8), (24, 264), (1000, 520), (17, 1024), (300, 2048 + 8)])
def test_gemv_plain(built_lib, n, k):
    W, x = rand_bf16(n, k, seed=n + k, scale=0.02), rand_bf16(1, 1, k, seed=1)
    got = ops.linear(x.to(DEV), W.to(DEV))
    want = orc.linear(x, W, None)
    assert got.shape == (1, 1, n)
    assert_close_bf16(got, want, 1, f"gemv {n}x{k}", atol=1e-4, frac_exact=0.98)


def test_gemv_bias_double_rounding(built_lib):
    s = models.QWEN25_05B
    n, k = s.q_dim + 2 * s.kv_dim, s.hidden
    W, x, b = rand_bf16(n, k, seed=2, scale=0.02), rand_bf16(1, 1, k, seed=3), rand_bf16(n, seed=4, scale=0.02)
    got = ops.linear(x.to(DEV), W.to(DEV), b.to(DEV))
    assert_close_bf16(got, orc.linear(x, W, b, three_d=True), 1, "gemv+bias", atol=1e-4, frac_exact=0.98)


def test_gemv_multi_row_and_lm_head_slice(built_lib):
    W, x = rand_bf16(4096, 896, seed=5, scale=0.02), rand_bf16(1, 3, 896, seed=6)
    got = ops.linear(x.to(DEV), W.to(DEV))
    assert_close_bf16(got, orc.linear(x, W, None), 1, "gemv m=3", atol=1e-4, frac_exact=0.98)


def test_gemv_linear_golden(built_lib):
    # TEST_Function.func_linear (test_function.cpp:234-248) padded to k = 8 (rows must be 16-byte multiples)
    x = torch.zeros(1, 2, 8)
    x[0, :, :3] = torch.tensor([[-0.3089, 0.5301, -0.0245], [1.5852, 0.8954, 0.7485]])
    W = torch.zeros(8, 8)
    W[:2, :3] = torch.tensor([[0.8397, 1.7990, -0.2738], [-0.8910, -0.6746, 0.3419]])
    b = torch.zeros(8)
    b[:2] = torch.tensor([-0.9601, -1.4163])
    got = ops.linear(x.to(torch.bfloat16).to(DEV), W.to(torch.bfloat16).to(DEV), b.to(torch.bfloat16).to(DEV))
    assert abs(float(got.float()[0, :, :2].sum()) - (-3.1661377)) < 5e-2  # bf16 storage of the fp32 golden


def test_gemv_rejects_bad_shapes(built_lib):
    with pytest.raises(B200Error):
        ops.linear(rand_bf16(1, 1, 12).to(DEV), rand_bf16(8, 12).to(DEV))  # k % 8 != 0
    with pytest.raises(B200Error):
        ops.linear(rand_bf16(1, 1, 16).to(DEV), rand_bf16(8, 8).to(DEV))  # mismatched k


@pytest.mark.parametrize("spec", [models.QWEN25_05B, models.LLAMA32_3B, models.QWEN3_17B, models.MISTRAL_7B],
                         ids=lambda s: s.name)
def test_fused_layer_pieces(built_lib, spec):
    """The four fused launches of a layer, each against the oracle's unfused sequence."""
    H, I = spec.hidden, spec.intermediate
    x = rand_bf16(H, seed=11)
    nw = (1 + rand_bf16(H, seed=12, scale=0.02).float()).to(torch.bfloat16)
    # 1. RMSNorm → qkv (+bias)
    n = spec.q_dim + 2 * spec.kv_dim
    Wq, bq = rand_bf16(n, H, seed=13, scale=0.02), rand_bf16(n, seed=14, scale=0.02)
    got = ops.gemv_fused(x.to(DEV), Wq.to(DEV), norm_weight=nw.to(DEV), eps=spec.rms_eps, bias=bq.to(DEV))
    want = orc.linear(orc.rms_norm(x, nw, spec.rms_eps), Wq, bq)
    assert_close_bf16(got, want, 2, "norm+qkv+bias", atol=2e-4, frac_exact=0.9)
    # 2. o_proj + residual (in place)
    a = rand_bf16(spec.q_dim, seed=15)
    Wo = rand_bf16(H, spec.q_dim, seed=16, scale=0.02)
    res = x.to(DEV).clone()
    got = ops.gemv_fused(a.to(DEV), Wo.to(DEV), residual=res)
    want = orc.add(x, orc.linear(a, Wo, None))
    assert_close_bf16(got, want, 2, "o_proj+residual", atol=1e-4, frac_exact=0.98)  # 1 ulp of acc can be 2 ulp of the sum
    # 3. RMSNorm → gate|up → SiLU·mul
    Wgu = rand_bf16(2 * I, H, seed=17, scale=0.02)
    got = ops.gemv_fused(x.to(DEV), Wgu.to(DEV), norm_weight=nw.to(DEV), eps=spec.rms_eps, silu_mul=True)
    want = orc.silu_mul(orc.linear(orc.rms_norm(x, nw, spec.rms_eps), Wgu, None))
    assert got.shape == (I,)
    assert_close_bf16(got, want, 3, "norm+gate_up+silu_mul", atol=2e-4, frac_exact=0.9)
    # 4. down_proj + residual
    m = rand_bf16(I, seed=18, scale=0.3)
    Wd = rand_bf16(H, I, seed=19, scale=0.02)
    got = ops.gemv_fused(m.to(DEV), Wd.to(DEV), residual=x.to(DEV))
    want = orc.add(x, orc.linear(m, Wd, None))
    assert_close_bf16(got, want, 2, "down+residual", atol=1e-4, frac_exact=0.98)


def test_gemv_linearity_full_size(built_lib):
    """Size-independent property at the largest decode shape (lm_head of Qwen2.5: 151936 × 896, 272 MB):
    W·(x1 + x2) == W·x1 + W·x2 up to bf16 rounding, and rows of a one-hot x reproduce W's columns exactly."""
    n, k = 151936, 896
    W = (torch.randn(n, k, device=DEV, generator=torch.Generator(DEV).manual_seed(0)) * 0.02).to(torch.bfloat16)
    onehot = torch.zeros(1, 1, k, dtype=torch.bfloat16, device=DEV)
    onehot[..., 123] = 1.0
    col = ops.linear(onehot, W)
    assert torch.equal(col.view(-1), W[:, 123]), "one-hot GEMV must return the weight column bit-exactly"
    x1 = torch.randn(1, 1, k, device=DEV).to(torch.bfloat16)
    y = ops.linear(x1, W).float()
    ref = (x1.float().view(1, k) @ W.float().t()).view(-1)  # torch fp32 matmul on the same device as a cross-check
    assert_close_bf16(y.view(-1), ref.to(torch.bfloat16).float(), 1, "lm_head GEMV vs fp32 matmul", atol=1e-4,
                      frac_exact=0.97)
