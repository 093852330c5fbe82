"""tcgen05 prefill GEMM against the oracle's Linear (fp32 accumulate, one bf16 rounding).  Tolerance ≤ 1 bf16 ulp (the
tensor core sums the fp32 products in a different order than torch's fp32 matmul), 1e-4 absolute near zero, ≥ 97 % of the
elements bit-identical; ragged M / N / K exercise TMA's zero fill."""
import pytest
import torch

from helpers import assert_close_bf16, orc
from tinygpt_b200 import ops
from tinygpt_b200._lib import B200Error

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rand_bf16(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 128, 256), (256, 384, 512), (16, 1152, 896), (200, 896, 4864),
                                   (1, 128, 64), (130, 136, 72), (2048, 4096, 2048)])
def test_gemm_vs_oracle(built_lib, M, N, K):
    a, w = rand_bf16(M, K, seed=M + K), rand_bf16(N, K, seed=N, scale=0.02)
    got = ops.gemm(a.to(DEV), w.to(DEV))
    assert got.shape == (M, N)
    if M * N * K <= 2 ** 31:
        want = orc.linear(a, w, None)
    else:  # big: fp32 matmul on the device as the cross-check
        want = (a.to(DEV).float() @ w.to(DEV).float().t()).to(torch.bfloat16).float().cpu()
    assert_close_bf16(got, want, 1, f"gemm {M}x{N}x{K}", atol=1e-4, frac_exact=0.97)


def test_gemm_matches_gemv_rows(built_lib):
    """Prefill and decode must agree: every row of the GEMM equals the decode GEMV of that row (≤ 1 ulp)."""
    a, w = rand_bf16(9, 896, seed=1), rand_bf16(1152, 896, seed=2, scale=0.02)
    big = ops.gemm(a.to(DEV), w.to(DEV)).float().cpu()
    rows = ops.linear(a.view(1, 9, 896).to(DEV), w.to(DEV)).float().cpu().view(9, 1152)
    assert_close_bf16(big, rows, 1, "gemm vs gemv", atol=1e-4, frac_exact=0.97)


def test_gemm_rejects_bad_shapes(built_lib):
    with pytest.raises(B200Error):
        ops.gemm(rand_bf16(4, 12).to(DEV), rand_bf16(8, 12).to(DEV))
