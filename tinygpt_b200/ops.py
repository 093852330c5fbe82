"""Host-side mirror of the TinyTorch functions on the decode path, over torch CUDA tensors → C ABI → sm_100a kernels.

Names and argument meaning follow tinytorch::function::* / op::* (third_party/TinyTorch/src/Function/FuncNNLayer.h,
FuncFused.h, FuncElemWise.h, FuncReduce.h); torch is used only for device memory and the current stream.
Every function raises B200Error when the CUDA library or an sm_100 device is missing — no fallback.
"""
from __future__ import annotations

from typing import Optional

import torch

from ._lib import B200Error, check, lib

BSHD, BHSD = 1, 0  # b200_qkv_layout


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _chk(t: torch.Tensor, name: str, dtype=torch.bfloat16) -> torch.Tensor:
    if not t.is_cuda:
        raise B200Error(f"{name}: tensor must live on a CUDA device (there is no CPU path)")
    if t.dtype != dtype:
        raise B200Error(f"{name}: expected dtype {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """function::linear(x, W, b) = op::matmul(x, W, false, true, b)  [ref: TT/Function/FuncNNLayer.h:14-18].
    x [..., k] bf16, W [n, k] bf16 → [..., n] bf16."""
    x, weight = _chk(x, "linear.x"), _chk(weight, "linear.weight")
    if bias is not None:
        bias = _chk(bias, "linear.bias")
    n, k = weight.shape
    if x.shape[-1] != k:
        raise B200Error(f"linear: x[..., {x.shape[-1]}] does not match W[{n}, {k}]")
    m = x.numel() // k
    y = torch.empty(*x.shape[:-1], n, dtype=torch.bfloat16, device=x.device)
    check(lib().b200_gemv_bf16(y.data_ptr(), x.data_ptr(), weight.data_ptr(), _ptr(bias), m, n, k, _stream()),
          "b200_gemv_bf16")
    return y


def gemm(a: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """Batched Linear without bias on the tcgen05 tensor cores: a [..., k] · W[n, k]ᵀ → [..., n] (prefill, m > 1)."""
    a, weight = _chk(a, "gemm.a"), _chk(weight, "gemm.weight")
    n, k = weight.shape
    if a.shape[-1] != k:
        raise B200Error(f"gemm: a[..., {a.shape[-1]}] does not match W[{n}, {k}]")
    m = a.numel() // k
    c = torch.empty(*a.shape[:-1], n, dtype=torch.bfloat16, device=a.device)
    check(lib().b200_gemm_bf16(c.data_ptr(), a.data_ptr(), weight.data_ptr(), m, n, k, _stream()), "b200_gemm_bf16")
    return c


def rms_norm(x: torch.Tensor, weight: Optional[torch.Tensor], eps: float) -> torch.Tensor:
    """function::rmsNorm(x, {dim}, w, eps)  [ref: TT/Function/FuncNNLayer.h:220-227]."""
    x = _chk(x, "rms_norm.x")
    if weight is not None:
        weight = _chk(weight, "rms_norm.weight")
    dim = x.shape[-1]
    y = torch.empty_like(x)
    check(lib().b200_rmsnorm_bf16(y.data_ptr(), x.data_ptr(), _ptr(weight), x.numel() // dim, dim, float(eps),
                                  _stream()), "b200_rmsnorm_bf16")
    return y


def rope_apply(x: torch.Tensor, table: torch.Tensor, offset: int, layout: int = BSHD) -> torch.Tensor:
    """function::ropeApply(x, rope, offset, layout)  [ref: TT/Function/FuncNNLayer.h:279-286]. x is 4-D."""
    x, table = _chk(x, "rope_apply.x"), _chk(table, "rope_apply.table", torch.float32)
    if x.dim() != 4:
        raise B200Error("rope_apply: input must be 4-D")
    if layout == BSHD:
        B, S, N, D = x.shape
    else:
        B, N, S, D = x.shape
    if table.shape[-2] != D or offset + S > table.shape[0]:
        raise B200Error("rope_apply: table does not cover the requested positions / head_dim")
    y = torch.empty_like(x)
    check(lib().b200_rope_bf16(y.data_ptr(), x.data_ptr(), table.data_ptr(), B, S, N, D, int(offset), layout,
                               _stream()), "b200_rope_bf16")
    return y


def rope_init(head_dim: int, ctx: int, theta: float, scaling=None, device="cuda") -> torch.Tensor:
    """op::ropeInit  [ref: TT/Operation/OpNNLayerCuda.cuh:621-656] → fp32 [ctx, head_dim, 2] built on the device."""
    t = torch.empty(ctx, head_dim, 2, dtype=torch.float32, device=device)
    f, hi, lo, orig = (0.0, 0.0, 0.0, 0) if scaling is None else (
        scaling.factor, scaling.high_freq_factor, scaling.low_freq_factor, scaling.original_context_length)
    check(lib().b200_rope_init_f32(t.data_ptr(), head_dim, ctx, float(theta), float(f), float(hi), float(lo),
                                   int(orig), _stream()), "b200_rope_init_f32")
    return t


def flash_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, is_causal: bool = False) -> torch.Tensor:
    """function::flashAttention(q, k, v, isCausal), BSHD  [ref: TT/Function/FuncNNLayer.h:270-277]."""
    q, k, v = _chk(q, "flash_attention.q"), _chk(k, "flash_attention.k"), _chk(v, "flash_attention.v")
    B, Sq, Hq, D = q.shape
    Skv, Hkv = k.shape[1], k.shape[2]
    o = torch.empty_like(q)
    check(lib().b200_attn_bf16(o.data_ptr(), q.data_ptr(), k.data_ptr(), v.data_ptr(), B, Sq, Skv, Hq, Hkv, D,
                               int(bool(is_causal)), _stream()), "b200_attn_bf16")
    return o


def silu_mul(x: torch.Tensor) -> torch.Tensor:
    """function::siluMul(x): last dim = [gate | up]  [ref: TT/Function/FuncFused.h:14-21]."""
    x = _chk(x, "silu_mul.x")
    I = x.shape[-1] // 2
    if x.shape[-1] % 2:
        raise B200Error("silu_mul: last dimension must be even")
    y = torch.empty(*x.shape[:-1], I, dtype=torch.bfloat16, device=x.device)
    check(lib().b200_silu_mul_bf16(y.data_ptr(), x.data_ptr(), x.numel() // (2 * I), I, _stream()),
          "b200_silu_mul_bf16")
    return y


def add(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """function::add(a, b) with alpha = 1, same shapes  [ref: TT/Function/FuncElemWise.h:74]."""
    a, b = _chk(a, "add.a"), _chk(b, "add.b")
    if a.shape != b.shape:
        raise B200Error("add: the decode path only adds same-shape tensors")
    y = torch.empty_like(a)
    check(lib().b200_add_bf16(y.data_ptr(), a.data_ptr(), b.data_ptr(), a.numel(), _stream()), "b200_add_bf16")
    return y


def embedding(table: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """function::embedding(ids, table) → op::indexAdvance  [ref: TT/Function/FuncNNLayer.h:202-207]; ids Int64."""
    table, ids = _chk(table, "embedding.table"), _chk(ids, "embedding.ids", torch.int64)
    V, H = table.shape
    y = torch.empty(*ids.shape, H, dtype=torch.bfloat16, device=table.device)
    if ids.numel():
        check(lib().b200_embedding_bf16(y.data_ptr(), table.data_ptr(), ids.data_ptr(), ids.numel(), V, H, _stream()),
              "b200_embedding_bf16")
    return y


def argmax(logits: torch.Tensor, keepdim: bool = True) -> torch.Tensor:
    """function::argmax(logits, -1, keepdim): fp32 compare, HIGHEST index wins ties (reference CUDA rule)
    [ref: src/engine/Sampler.cpp:23-29; TT/Operation/OpReduceCuda.cuh:145-156]."""
    logits = _chk(logits, "argmax.logits")
    V = logits.shape[-1]
    rows = logits.numel() // V
    ws = torch.zeros(lib().b200_argmax_workspace_bytes(rows, V), dtype=torch.uint8, device=logits.device)
    out = torch.empty(rows, dtype=torch.int64, device=logits.device)
    check(lib().b200_argmax_bf16(out.data_ptr(), logits.data_ptr(), rows, V, ws.data_ptr(), _stream()),
          "b200_argmax_bf16")
    shape = list(logits.shape[:-1]) + ([1] if keepdim else [])
    return out.view(shape)


# ---- fused building blocks (the kernels the engine actually runs)
def gemv_fused(x, weight, *, norm_weight=None, eps=0.0, bias=None, residual=None, silu_mul=False) -> torch.Tensor:
    """[RMSNorm] → GEMV → {bias | residual | SiLU·mul over merged gate|up} in ONE launch (m = 1)."""
    x, weight = _chk(x, "gemv_fused.x"), _chk(weight, "gemv_fused.weight")
    rows, k = weight.shape
    nseg = 2 if silu_mul else 1
    n = rows // nseg
    y = torch.empty(n, dtype=torch.bfloat16, device=x.device)
    check(lib().b200_gemv_fused_bf16(y.data_ptr(), x.data_ptr(), weight.data_ptr(), n, k, nseg, _ptr(norm_weight),
                                     float(eps), _ptr(bias), _ptr(residual), int(silu_mul), _stream()),
          "b200_gemv_fused_bf16")
    return y


def attn_decode(qkv, kcache, vcache, *, q_heads, kv_heads, head_dim, pos=None, fixed_len=0, rope_table=None,
                q_norm=None, k_norm=None, eps=0.0) -> torch.Tensor:
    """Fused decode attention of one layer (see include/b200_decode.h b200_attn_decode_bf16)."""
    qkv, kcache, vcache = _chk(qkv, "attn_decode.qkv"), _chk(kcache, "attn_decode.kcache"), _chk(vcache, "attn_decode.vcache")
    max_ctx = kcache.shape[0]
    ws = torch.zeros(lib().b200_attn_decode_workspace_bytes(q_heads, kv_heads, head_dim, max_ctx), dtype=torch.uint8,
                     device=qkv.device)
    out = torch.empty(q_heads * head_dim, dtype=torch.bfloat16, device=qkv.device)
    check(lib().b200_attn_decode_bf16(out.data_ptr(), qkv.data_ptr(), _ptr(q_norm), _ptr(k_norm), float(eps),
                                      _ptr(rope_table), _ptr(pos), int(fixed_len), kcache.data_ptr(),
                                      vcache.data_ptr(), q_heads, kv_heads, head_dim, max_ctx, ws.data_ptr(),
                                      _stream()), "b200_attn_decode_bf16")
    return out


def sample(logits: torch.Tensor, temperature: float, top_k: int = 0, top_p: float = 1.0, min_p: float = 0.0,
           u: float = 0.5) -> torch.Tensor:
    """Sampler::sample for one row of bf16 logits [V] with the uniform number u supplied by the caller
    [ref: src/engine/Sampler.cpp:31-78].  Returns an int64 device tensor [1]."""
    logits = _chk(logits, "sample.logits").view(-1)
    ws = torch.empty(lib().b200_sample_workspace_bytes(), dtype=torch.uint8, device=logits.device)
    out = torch.empty(1, dtype=torch.int64, device=logits.device)
    check(lib().b200_sample_bf16(out.data_ptr(), logits.data_ptr(), logits.numel(), float(temperature), int(top_k),
                                 float(top_p), float(min_p), float(u), ws.data_ptr(), _stream()), "b200_sample_bf16")
    return out
