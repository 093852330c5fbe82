"""Per-op parity of the sm_100a kernels (through the C ABI) against the CPU oracle — boundary B of SURVEY.md §8b.

Tolerances (stated per test): exact for copy/compare ops (embedding, add, argmax, silu_mul given identical inputs up
to expf ulp), ≤ 1 bf16 ulp where fp32 summation order or rsqrt/exp implementation differs from the oracle's."""
import math

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


# ------------------------------------------------------------------------------------------------- rmsnorm
@pytest.mark.parametrize("rows,dim", [(1, 896), (1, 3072), (1, 2048), (1, 4096), (24, 128), (3, 8), (5, 1000)])
def test_rmsnorm(built_lib, rows, dim):
    x, w = rand_bf16(rows, dim, seed=1), (1 + rand_bf16(dim, seed=2, scale=0.02).float()).to(torch.bfloat16)
    got = ops.rms_norm(x.to(DEV), w.to(DEV), 1e-6)
    want = orc.rms_norm(x, w, 1e-6)
    # rsqrtf (MUFU) vs 1/sqrt and the reduction order may move a value across a rounding boundary: ≤ 1 ulp
    assert_close_bf16(got, want, 1, f"rmsnorm {rows}x{dim}", frac_exact=0.97)


def test_rmsnorm_no_weight_and_golden(built_lib):
    # reference golden vector TEST_Function.func_rmsNorm (test_function.cpp:385-391), here through bf16 storage
    x = torch.tensor([[1.4176, 0.1874, 0.8367], [-0.1203, 2.5638, -1.2554]]).to(torch.bfloat16)
    w = torch.tensor([1.4072, -0.4768, -0.6006]).to(torch.bfloat16)
    got = ops.rms_norm(x.to(DEV), w.to(DEV), 1e-8).float().cpu()
    want = torch.tensor([[2.0855, -0.0934, -0.5254], [-0.1026, -0.7410, 0.4571]])
    assert torch.allclose(got, want, atol=2e-2), got  # bf16 inputs: 3 significant digits
    got2 = ops.rms_norm(x.to(DEV), None, 1e-8)
    assert_close_bf16(got2, orc.rms_norm(x, None, 1e-8), 1, "rmsnorm no weight")


# ---------------------------------------------------------------------------------------------------- rope
@pytest.mark.parametrize("layout", ["BSHD", "BHSD"])
@pytest.mark.parametrize("B,S,N,D,offset", [(1, 1, 14, 64, 17), (1, 1, 8, 128, 130), (2, 5, 3, 64, 0), (1, 33, 4, 128, 7)])
def test_rope(built_lib, layout, B, S, N, D, offset):
    table = orc.rope_table(D, 256, 1e6)
    shape = (B, S, N, D) if layout == "BSHD" else (B, N, S, D)
    x = rand_bf16(*shape, seed=3)
    got = ops.rope_apply(x.to(DEV), table.to(DEV), offset, ops.BSHD if layout == "BSHD" else ops.BHSD)
    want = orc.rope_apply(x, table, offset, layout)
    # fp32 x1*c - x2*s: the device contracts to an FMA, the oracle does not: ≤ 1 ulp, nearly all identical
    assert_close_bf16(got, want, 1, f"rope {layout} {shape}", frac_exact=0.99)


def test_rope_table_device_matches_host(built_lib):
    for spec in (models.QWEN25_05B.with_ctx(512), models.LLAMA32_3B.with_ctx(512)):
        host = models.rope_table(spec)
        dev = ops.rope_init(spec.head_dim, spec.max_ctx, spec.rope_theta, spec.rope_scaling).cpu()
        # device powf/cosf/sinf vs numpy float32: a few fp32 ulps at large angles (|angle| up to 511 rad)
        assert torch.allclose(host, dev, atol=2e-4, rtol=0), float((host - dev).abs().max())
        assert float((host - dev).abs().mean()) < 5e-6


def test_rope_errors(built_lib):
    table = orc.rope_table(64, 16, 1e4).to(DEV)
    with pytest.raises(B200Error):
        ops.rope_apply(rand_bf16(1, 20, 2, 64).to(DEV), table, 0)  # positions beyond the table
    with pytest.raises(B200Error):
        ops.rope_apply(rand_bf16(1, 2, 64).to(DEV), table, 0)  # not 4-D


# ---------------------------------------------------------------------------------------------- silu / add
@pytest.mark.parametrize("rows,I", [(1, 4864), (1, 8192), (1, 14336), (3, 40), (1, 1)])
def test_silu_mul(built_lib, rows, I):
    gu = rand_bf16(rows, 2 * I, seed=4, scale=2.0)
    got = ops.silu_mul(gu.to(DEV))
    want = orc.silu_mul(gu)
    # device expf vs torch exp may differ in the last fp32 bit: ≤ 1 ulp after the two bf16 roundings
    assert_close_bf16(got, want, 1, f"silu_mul {rows}x{I}", frac_exact=0.995)


def test_silu_golden(built_lib):
    # TEST_Function.func_silu (test_function.cpp:139-145): silu(x) * 1
    x = torch.tensor([-1.0, -0.5, 0.5, 1.0])
    gu = torch.cat([x, torch.ones(4)]).to(torch.bfloat16).view(1, 8)
    got = ops.silu_mul(gu.to(DEV)).float().cpu().view(-1)
    assert torch.allclose(got, torch.tensor([-0.2689, -0.1888, 0.3112, 0.7311]), atol=2e-3)


@pytest.mark.parametrize("n", [1, 7, 896, 4096, 100003])
def test_add_exact(built_lib, n):
    a, b = rand_bf16(n, seed=5), rand_bf16(n, seed=6, scale=0.1)
    got = ops.add(a.to(DEV), b.to(DEV))
    want = orc.add(a, b)
    assert torch.equal(got.float().cpu(), want), "bf16 add must be bit-exact"


# ----------------------------------------------------------------------------------------------- embedding
def test_embedding_exact(built_lib):
    table = rand_bf16(1000, 896, seed=7)
    ids = torch.tensor([[0, 999, 5, 5, 123]], dtype=torch.int64)
    got = ops.embedding(table.to(DEV), ids.to(DEV))
    assert got.shape == (1, 5, 896)
    assert torch.equal(got.cpu(), table[ids])
    odd = rand_bf16(10, 12, seed=8)  # H not a multiple of 8 → scalar path
    assert torch.equal(ops.embedding(odd.to(DEV), ids.clamp(max=9).to(DEV)).cpu(), odd[ids.clamp(max=9)])
    empty = ops.embedding(table.to(DEV), torch.empty(1, 0, dtype=torch.int64, device=DEV))
    assert empty.shape == (1, 0, 896)
    with pytest.raises(B200Error):
        ops.embedding(table.to(DEV), ids.to(torch.int32).to(DEV))  # reference asserts Int64 (FuncNNLayer.h:205)


# -------------------------------------------------------------------------------------------------- argmax
@pytest.mark.parametrize("V", [32768, 128256, 151936, 50257, 7, 1])
def test_argmax_exact(built_lib, V):
    lg = rand_bf16(1, V, seed=9)
    got = ops.argmax(lg.to(DEV))
    assert got.shape == (1, 1)
    assert int(got) == int(orc.argmax_last(lg))


def test_argmax_tie_rule_last_index_wins(built_lib):
    V = 151936
    lg = torch.zeros(3, V, dtype=torch.bfloat16)
    lg[0, [5, 70000, 151935]] = 3.0          # three equal maxima → highest index
    lg[1, [0, 1]] = 1.0
    lg[2, :] = -1.0                           # everything equal → V - 1
    got = ops.argmax(lg.to(DEV), keepdim=False).cpu()
    assert got.tolist() == [151935, 1, V - 1]
    assert got.tolist() == orc.argmax_last(lg).tolist()
    # first-max rule (torch / reference CPU path) would give [5, 0, 0]: make sure we do NOT follow it
    assert got.tolist() != torch.argmax(lg.float(), dim=-1).tolist()


def test_argmax_repeatable_ticket_reset(built_lib):
    lg = rand_bf16(4, 32768, seed=10).to(DEV)
    a = ops.argmax(lg).cpu()
    b = ops.argmax(lg).cpu()
    assert torch.equal(a, b) and torch.equal(a.view(-1), orc.argmax_last(lg.cpu()))


# ------------------------------------------------------------------------------------------------ CPU tensors
def test_no_cpu_path(built_lib):
    with pytest.raises(B200Error):
        ops.add(rand_bf16(8), rand_bf16(8))
    with pytest.raises(B200Error):
        ops.linear(rand_bf16(1, 8), rand_bf16(8, 8))
