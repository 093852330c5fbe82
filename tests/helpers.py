"""Shared test helpers: oracle access, bf16 ulp metrics, spec conversion."""
from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from oracle import decode_oracle as orc  # noqa: E402  (tests are allowed to import the oracle)
from tinygpt_b200 import models  # noqa: E402


def to_oracle_cfg(spec: models.ModelSpec) -> orc.ModelConfig:
    sc = spec.rope_scaling
    return orc.ModelConfig(spec.name, spec.hidden, spec.layers, spec.q_heads, spec.kv_heads, spec.head_dim,
                           spec.intermediate, spec.vocab, spec.rope_theta, spec.rms_eps, spec.tie, spec.qkv_bias,
                           spec.qk_norm, spec.max_ctx,
                           None if sc is None else orc.RopeScaling(sc.factor, sc.high_freq_factor, sc.low_freq_factor,
                                                                   sc.original_context_length))


def bf16_ulp_diff(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """|a - b| in units of bf16 ulps at max(|a|,|b|) (both given as float tensors holding bf16 values)."""
    a, b = a.float().cpu(), b.float().cpu()
    mag = torch.maximum(a.abs(), b.abs()).clamp_min(2.0 ** -126)
    ulp = torch.exp2(torch.floor(torch.log2(mag)) - 7)
    return (a - b).abs() / ulp


def assert_close_bf16(got: torch.Tensor, want: torch.Tensor, max_ulp: float, what: str, atol: float = 0.0,
                      frac_exact: float | None = None):
    """got/want hold bf16 values.  Passes when every element is within max_ulp bf16 ulps OR within atol absolute
    (cancellation near zero makes ulps meaningless there)."""
    got_f, want_f = got.float().cpu(), want.float().cpu()
    assert got_f.shape == want_f.shape, f"{what}: shape {tuple(got_f.shape)} vs {tuple(want_f.shape)}"
    assert torch.isfinite(got_f).all(), f"{what}: non-finite output"
    d = bf16_ulp_diff(got_f, want_f)
    bad = (d > max_ulp) & ((got_f - want_f).abs() > atol)
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{bad.numel()} elements differ by more than {max_ulp} bf16 ulp "
                           f"(max {float(d.max()):.2f} ulp, max abs {float((got_f - want_f).abs().max()):.3e})")
    if frac_exact is not None:
        exact = float((got_f == want_f).float().mean())
        assert exact >= frac_exact, f"{what}: only {exact:.4f} of elements bit-identical (< {frac_exact})"
