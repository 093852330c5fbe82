"""CPU oracle of TinyGPT's decode hot path — TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, the arithmetic of the reference's CUDA path for one token of a Llama-family model
(keith2018/TinyGPT @ e3b6ab1 with TinyTorch @ ab62352 and TinyFA @ 4e18516).  It exists to CHECK the sm_100a kernels
in tinygpt_b200/csrc; nothing in the product path may import it (only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs do).  The product path has no CPU fallback.

Pinning status (see DESIGN.md "Oracle"):
  * pinned against the reference's own golden vectors (tests/golden/reference_vectors.json: func_rmsNorm, func_silu,
    func_linear, func_sdpAttention, TEST_Module.rope) — tests/test_oracle_golden.py;
  * pinned against outputs of the reference itself, compiled from /root/reference by oracle/Makefile into
    oracle/_ref and run on the CPU in fp32 (per-op and whole-model logits; fixtures in tests/golden/ref_cpu_*.npz);
  * the bf16 rounding points follow the reference's CUDA sources by reading (cited per function); the reference's
    CUDA build cannot run in the build container (no GPU) and its sources cannot travel to the GPU box, and cuBLAS'
    summation order is not reproducible — so bf16 parity is "within tolerance", never bit-exact, except where the
    arithmetic is exact (embedding, add, argmax, KV append).

`dtype="bf16"` applies every rounding the CUDA path applies; `dtype="fp32"` disables them (for comparison with the
reference's fp32 CPU build).

File:line citations are relative to the reference checkout; TT/ = third_party/TinyTorch/src/,
TFA/ = third_party/TinyTorch/third_party/TinyFA/csrc/flash_attn/.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch

F32 = torch.float32
BF16 = torch.bfloat16


def rnd(x: torch.Tensor, dtype: str) -> torch.Tensor:
    """The reference's `static_cast<T>(float)` rounding point: RNE to bf16 (TT/Utils/BFloat16.h:72-96), kept as fp32."""
    if dtype == "bf16":
        return x.to(BF16).to(F32)
    return x.to(F32)


# ----------------------------------------------------------------------------------------------------------- linear
def linear(x: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor], dtype: str = "bf16",
           three_d: bool = True) -> torch.Tensor:
    """op::matmul(x, W, false, true, bias).

    CUDA bf16: bf16×bf16 products accumulated in fp32, ONE rounding of the accumulator to bf16
    (TT/Operation/OpLinalgCuda.cuh:276-293, CUDA_R_32F compute).  With a bias and a 3-D input (every Linear of the
    decode path sees [B,S,H]) the bias is added by a separate elementwise kernel on the already-rounded result —
    a second rounding (TT/Operation/OpLinalg.cpp:273-275).  2-D inputs pre-fill C with the bias and use beta = 1
    (single rounding, TT/Operation/OpLinalgCuda.cuh:141-146,193-216).
    """
    acc = x.to(F32) @ W.to(F32).t()
    if bias is None:
        return rnd(acc, dtype)
    if three_d:
        return rnd(rnd(acc, dtype) + bias.to(F32), dtype)
    return rnd(acc + bias.to(F32), dtype)


# ---------------------------------------------------------------------------------------------------------- rmsnorm
def rms_norm(x: torch.Tensor, w: Optional[torch.Tensor], eps: float, dtype: str = "bf16") -> torch.Tensor:
    """kNormSmall/kNormLarge<RMSNorm> (TT/Operation/OpNNLayerCuda.cuh:252-357): fp32 Σx², inv = rsqrt(mean + eps),
    normed = x*inv; normed *= w (fp32); one rounding.  (HF rounds before the weight multiply; the reference does not.)"""
    xf = x.to(F32)
    stat = (xf * xf).sum(dim=-1, keepdim=True) / xf.shape[-1]
    inv = 1.0 / torch.sqrt(stat + torch.tensor(eps, dtype=F32))
    y = xf * inv
    if w is not None:
        y = y * w.to(F32)
    return rnd(y, dtype)


# ------------------------------------------------------------------------------------------------------------- rope
@dataclass
class RopeScaling:
    """RopeScalingConfig {factor, highFreqFactor, lowFreqFactor, originalContextLength} (TT/Operation/OpNNLayer.h:15-21)."""
    factor: float
    high_freq_factor: float
    low_freq_factor: float
    original_context_length: int


def rope_table(head_dim: int, ctx: int, theta: float, scaling: Optional[RopeScaling] = None) -> torch.Tensor:
    """ropeInit (TT/Operation/OpNNLayerCuda.cuh:359-410, host :621-656): fp32 table [ctx, head_dim, 2] = (cos, sin),
    duplicated for both halves.  invFreq_i = 1 / powf(theta, (2i)/head_dim); llama3 scaling per kRopeApplyScaling."""
    half = head_dim // 2
    i = np.arange(half, dtype=np.float32)
    expo = (i * np.float32(2.0)) / np.float32(head_dim)
    inv = (np.float32(1.0) / np.power(np.float32(theta), expo, dtype=np.float32)).astype(np.float32)
    if scaling is not None:
        out = inv.copy()
        orig = np.float32(scaling.original_context_length)
        low_wave = orig / np.float32(scaling.low_freq_factor)
        high_wave = orig / np.float32(scaling.high_freq_factor)
        for j in range(half):
            f = inv[j]
            wave = np.float32(2.0) * np.float32(math.pi) / f
            if wave > low_wave:
                out[j] = f / np.float32(scaling.factor)
            elif wave < high_wave:
                pass
            else:
                smooth = (orig / wave - np.float32(scaling.low_freq_factor)) / (
                    np.float32(scaling.high_freq_factor) - np.float32(scaling.low_freq_factor))
                scaled = f / np.float32(scaling.factor)
                out[j] = (np.float32(1.0) - smooth) * scaled + smooth * f
        inv = out.astype(np.float32)
    pos = np.arange(ctx, dtype=np.float32)[:, None]
    ang = (pos * inv[None, :]).astype(np.float32)
    c = np.cos(ang, dtype=np.float32)
    s = np.sin(ang, dtype=np.float32)
    tab = np.empty((ctx, head_dim, 2), dtype=np.float32)
    tab[:, :half, 0] = c
    tab[:, :half, 1] = s
    tab[:, half:, 0] = c
    tab[:, half:, 1] = s
    return torch.from_numpy(tab)


def rope_apply(x: torch.Tensor, table: torch.Tensor, offset: int, layout: str = "BSHD",
               dtype: str = "bf16") -> torch.Tensor:
    """kRopeApply (TT/Operation/OpNNLayerCuda.cuh:412-440): rotate-half, fp32 math, position = offset + t,
    y[i] = x1*c - x2*s ; y[i+half] = x2*c + x1*s ; one rounding each.  layout BSHD ([B,S,N,D]) or BHSD ([B,N,S,D])."""
    xf = x.to(F32)
    if layout == "BHSD":
        xf = xf.transpose(1, 2)
    B, S, N, D = xf.shape
    half = D // 2
    rows = table[offset:offset + S]  # [S, D, 2]
    c = rows[:, :half, 0].view(1, S, 1, half)
    s = rows[:, :half, 1].view(1, S, 1, half)
    x1, x2 = xf[..., :half], xf[..., half:]
    y = torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], dim=-1)
    if layout == "BHSD":
        y = y.transpose(1, 2)
    return rnd(y.contiguous(), dtype)


# -------------------------------------------------------------------------------------------------------- attention
_TFA_BC = {64: 128, 128: 64}  # SM8x bf16 configs: hd64 → Br128/Bc128, hd128 → Br128/Bc64 (TFA/config.cuh:261-268)
_INV_SQRT = {32: 0.17677669529663689, 64: 0.125, 96: 0.10206207261596576, 128: 0.08838834764831845,
             192: 0.07216878364870323, 256: 0.0625}
_LOG2E = 1.4426950408889634


def flash_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, causal: bool, dtype: str = "bf16",
                    model_p_rounding: bool = True) -> torch.Tensor:
    """tfa::flashAttn, BSHD, GQA (TFA/mma/kernel.cuh:18-203, TFA/mma/softmax.cuh:67-131).

    KV tiles of Bc columns are visited LAST → FIRST; per tile: S = q·kᵀ (fp32 accumulate), running max, rescale by
    exp2((old-new)·scale), P = exp2(S·scale − max·scale) with scale = log2(e)/sqrt(hd) (fp32 constant,
    TFA/utils.cuh:69-74), row sum of the fp32 P, then P is rounded to bf16 for O += P·V (TFA/mma/layout.cuh:88-97);
    O·(1/rowSum) rounded to bf16 (TFA/mma/memory.cuh:84-97).  Causal mask is top-left aligned: col > row masked
    (softmax.cuh:133-136), so decode (Sq = 1 against a longer cache) must be called non-causal, as the reference
    does (src/layer/Attention.h:108-109).  Masked positions beyond Skv never contribute.
    """
    B, Sq, Hq, D = q.shape
    Skv, Hkv = k.shape[1], k.shape[2]
    G = Hq // Hkv
    bc = _TFA_BC.get(D, 64)
    scale = np.float32(np.float32(_INV_SQRT.get(D, 1.0 / math.sqrt(D))) * np.float32(_LOG2E))
    scale_t = torch.tensor(float(scale), dtype=F32)
    qf = q.to(F32).permute(0, 2, 1, 3)                                   # [B,Hq,Sq,D]
    kf = k.to(F32).permute(0, 2, 1, 3).repeat_interleave(G, dim=1)       # [B,Hq,Skv,D]
    vf = v.to(F32).permute(0, 2, 1, 3).repeat_interleave(G, dim=1)
    m = torch.full((B, Hq, Sq), -math.inf, dtype=F32)
    l = torch.zeros((B, Hq, Sq), dtype=F32)
    o = torch.zeros((B, Hq, Sq, D), dtype=F32)
    rows = torch.arange(Sq).view(1, 1, Sq, 1)
    ntiles = (Skv + bc - 1) // bc
    for t in range(ntiles - 1, -1, -1):
        c0, c1 = t * bc, min(Skv, (t + 1) * bc)
        s = qf @ kf[:, :, c0:c1].transpose(-1, -2)                        # [B,Hq,Sq,bc]
        if causal:
            cols = torch.arange(c0, c1).view(1, 1, 1, -1)
            s = s.masked_fill(cols > rows, -math.inf)
        cur = s.max(dim=-1).values
        new_m = torch.maximum(m, cur)
        alpha = torch.where(torch.isinf(m) & (m < 0), torch.zeros_like(m), torch.exp2((m - new_m) * scale_t))
        alpha = torch.where(torch.isinf(new_m) & (new_m < 0), torch.zeros_like(alpha), alpha)
        max_scaled = torch.where(torch.isinf(new_m) & (new_m < 0), torch.zeros_like(new_m), new_m * scale_t)
        p = torch.exp2(s * scale_t - max_scaled.unsqueeze(-1))
        l = l * alpha + p.sum(dim=-1)
        o = o * alpha.unsqueeze(-1)
        pp = rnd(p, dtype) if model_p_rounding else p
        o = o + pp @ vf[:, :, c0:c1]
        m = new_m
    inv = torch.where(l > 0, 1.0 / l, torch.zeros_like(l))
    o = o * inv.unsqueeze(-1)
    return rnd(o.permute(0, 2, 1, 3).contiguous(), dtype)


def naive_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, causal: bool) -> torch.Tensor:
    """fp32 attention of the reference's test oracle (TFAroot/tests/cpp/cpu_reference.h:14-66); BSHD, GQA."""
    B, Sq, Hq, D = q.shape
    Skv, Hkv = k.shape[1], k.shape[2]
    G = Hq // Hkv
    qf = q.to(F32).permute(0, 2, 1, 3)
    kf = k.to(F32).permute(0, 2, 1, 3).repeat_interleave(G, dim=1)
    vf = v.to(F32).permute(0, 2, 1, 3).repeat_interleave(G, dim=1)
    s = (qf @ kf.transpose(-1, -2)) * (1.0 / math.sqrt(D))
    if causal:
        rows = torch.arange(Sq).view(1, 1, Sq, 1)
        cols = torch.arange(Skv).view(1, 1, 1, Skv)
        s = s.masked_fill(cols > rows, -math.inf)
    p = torch.softmax(s, dim=-1)
    return (p @ vf).permute(0, 2, 1, 3).contiguous()


# ----------------------------------------------------------------------------------------------- small elementwise
def silu(x: torch.Tensor, dtype: str = "bf16") -> torch.Tensor:
    """OpCudaSilu (TT/Operation/OpElemWiseCuda.cuh:124-131): fa / (1 + expf(-fa)), one rounding."""
    xf = x.to(F32)
    return rnd(xf / (1.0 + torch.exp(-xf)), dtype)


def silu_mul(gate_up: torch.Tensor, dtype: str = "bf16") -> torch.Tensor:
    """kSiluMul (TT/Operation/OpFusedCuda.cuh:15-29): last dim = [gate | up]; silu is rounded to T, THEN multiplied
    by up in T (second rounding)."""
    I = gate_up.shape[-1] // 2
    g, u = gate_up[..., :I], gate_up[..., I:]
    return rnd(silu(g, dtype) * u.to(F32), dtype)


def add(a: torch.Tensor, b: torch.Tensor, dtype: str = "bf16") -> torch.Tensor:
    """OpCudaAdd with alpha = 1 (TT/Operation/OpElemWiseCuda.cuh:133-144): a + b in T."""
    return rnd(a.to(F32) + b.to(F32), dtype)


def embedding(table: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """kIndex row gather (TT/Operation/OpTransformCuda.cuh:108-120): exact copy."""
    return table[ids.long()].to(F32)


def argmax_last(logits: torch.Tensor) -> torch.Tensor:
    """argmax over the last dim with the reference CUDA tie rule — the HIGHEST index among equal maxima
    (cudaWarpReduceIdx takes the other lane when max(other, val) == other, TT/Operation/OpReduceCuda.cuh:145-156)."""
    lf = logits.to(F32)
    V = lf.shape[-1]
    rev = torch.flip(lf, dims=[-1])
    return (V - 1 - torch.argmax(rev, dim=-1)).to(torch.int64)


# ------------------------------------------------------------------------------------------------------- the model
@dataclass
class ModelConfig:
    """The constants the reference parses from config.json (src/huggingface/ModelConfig.cpp:73-122)."""
    name: str
    hidden: int
    layers: int
    q_heads: int
    kv_heads: int
    head_dim: int
    intermediate: int
    vocab: int
    rope_theta: float
    rms_eps: float
    tie: bool = True
    qkv_bias: bool = False
    qk_norm: bool = False
    max_ctx: int = 4096
    rope_scaling: Optional[RopeScaling] = None

    @property
    def q_dim(self) -> int:
        return self.q_heads * self.head_dim

    @property
    def kv_dim(self) -> int:
        return self.kv_heads * self.head_dim


@dataclass
class KVCache:
    """KVCacheManager (src/engine/CacheManager.h:13-55): per layer K,V [1, ctx, Hkv, hd], append = concat on dim 1."""
    k: list = field(default_factory=list)
    v: list = field(default_factory=list)

    def past_length(self, layer: int) -> int:
        return 0 if layer >= len(self.k) or self.k[layer] is None else self.k[layer].shape[1]

    def append(self, layer: int, k: torch.Tensor, v: torch.Tensor):
        while len(self.k) <= layer:
            self.k.append(None)
            self.v.append(None)
        past = self.past_length(layer)
        if self.k[layer] is None:
            self.k[layer], self.v[layer] = k, v
        else:
            self.k[layer] = torch.cat([self.k[layer], k], dim=1)
            self.v[layer] = torch.cat([self.v[layer], v], dim=1)
        return self.k[layer], self.v[layer], past


def attention_block(cfg: ModelConfig, w: dict, layer: int, x: torch.Tensor, cache: KVCache, table: torch.Tensor,
                    dtype: str, trace: Optional[dict] = None) -> torch.Tensor:
    """Attention::forward (+ AttentionWithQKNorm::projectQKV) — src/layer/Attention.h:71-112,156-163."""
    p = f"model.layers.{layer}.self_attn."
    B, S, _ = x.shape
    qkv = linear(x, w[p + "qkv_proj.weight"], w.get(p + "qkv_proj.bias"), dtype)       # MergedLinear [q|k|v]
    q = qkv[..., :cfg.q_dim].reshape(B, S, cfg.q_heads, cfg.head_dim)
    k = qkv[..., cfg.q_dim:cfg.q_dim + cfg.kv_dim].reshape(B, S, cfg.kv_heads, cfg.head_dim)
    v = qkv[..., cfg.q_dim + cfg.kv_dim:].reshape(B, S, cfg.kv_heads, cfg.head_dim)
    if cfg.qk_norm:
        q = rms_norm(q, w[p + "q_norm.weight"], cfg.rms_eps, dtype)
        k = rms_norm(k, w[p + "k_norm.weight"], cfg.rms_eps, dtype)
    past = cache.past_length(layer)
    q = rope_apply(q, table, past, "BSHD", dtype)
    k = rope_apply(k, table, past, "BSHD", dtype)
    K, V, past_len = cache.append(layer, k, v)
    causal = past_len == 0                                                            # Attention.h:108
    if dtype == "bf16":
        o = flash_attention(q, K, V, causal, dtype)
    else:
        o = naive_attention(q, K, V, causal)
    if trace is not None:
        trace[f"l{layer}.qkv"] = qkv
        trace[f"l{layer}.attn"] = o
    o = o.reshape(B, S, cfg.q_dim)
    return linear(o, w[p + "o_proj.weight"], None, dtype)


def mlp_block(cfg: ModelConfig, w: dict, layer: int, x: torch.Tensor, dtype: str) -> torch.Tensor:
    """GatedMLP::forward — src/layer/GatedMLP.h:37-41: down(siluMul(gate_up(x)))."""
    p = f"model.layers.{layer}.mlp."
    gu = linear(x, w[p + "gate_up_proj.weight"], None, dtype)
    return linear(silu_mul(gu, dtype), w[p + "down_proj.weight"], None, dtype)


def forward(cfg: ModelConfig, w: dict, ids: torch.Tensor, cache: KVCache, table: torch.Tensor, dtype: str = "bf16",
            trace: Optional[dict] = None) -> torch.Tensor:
    """CausalLM::forward — src/model/GPTModel.h:51-58; DecoderLayer::forward — src/layer/DecoderLayer.h:38-43.
    ids [B,S] int64 → logits [B,S,V] (lm_head over every position, like the reference)."""
    x = embedding(w["model.embed_tokens.weight"], ids)
    for l in range(cfg.layers):
        p = f"model.layers.{l}."
        h = rms_norm(x, w[p + "input_layernorm.weight"], cfg.rms_eps, dtype)
        x = add(x, attention_block(cfg, w, l, h, cache, table, dtype, trace), dtype)
        h = rms_norm(x, w[p + "post_attention_layernorm.weight"], cfg.rms_eps, dtype)
        x = add(x, mlp_block(cfg, w, l, h, dtype), dtype)
        if trace is not None:
            trace[f"l{l}.x"] = x
    x = rms_norm(x, w["model.norm.weight"], cfg.rms_eps, dtype)
    head = w["model.embed_tokens.weight"] if cfg.tie else w["lm_head.weight"]
    return linear(x, head, None, dtype)


def generate_greedy(cfg: ModelConfig, w: dict, prompt: torch.Tensor, max_new: int, table: torch.Tensor,
                    dtype: str = "bf16", forced: Optional[torch.Tensor] = None):
    """GPTEngine::generateSync with the greedy sampler (src/engine/GPTEngine.cpp:154-174, Sampler.cpp:23-29).

    Returns (tokens [max_new], logits [max_new, V]).  `forced` (teacher forcing): feed these tokens instead of the
    oracle's own argmax, so that one near-tie cannot cascade when comparing against another implementation."""
    cache = KVCache()
    toks, logs = [], []
    logits = forward(cfg, w, prompt.view(1, -1), cache, table, dtype)[:, -1]
    for i in range(max_new):
        logs.append(logits[0])
        t = argmax_last(logits)
        toks.append(int(t[0]))
        if i == max_new - 1:
            break
        nxt = t if forced is None else forced[i].view(1)
        logits = forward(cfg, w, nxt.view(1, 1), cache, table, dtype)[:, -1]
    return torch.tensor(toks, dtype=torch.int64), torch.stack(logs)
