"""Model constants and synthetic checkpoints for the decode path.

Mirrors what the reference derives from config.json (src/huggingface/ModelConfig.cpp:73-122) and the weight-name /
merged-weight layout of its model classes (src/model/GPTModel.h:43-48, src/layer/Attention.h:61-68,
src/layer/GatedMLP.h:44-50, src/layer/Linear.h:64-79).  No weights are available offline, so checkpoints are seeded
synthetic tensors of the real shapes.
"""
from __future__ import annotations

import json
import struct
from dataclasses import dataclass, replace
from pathlib import Path
from typing import Dict, Optional

import numpy as np
import torch


@dataclass(frozen=True)
class RopeScaling:
    factor: float
    high_freq_factor: float
    low_freq_factor: float
    original_context_length: int


@dataclass(frozen=True)
class ModelSpec:
    name: str
    model_type: str          # llama | qwen2 | qwen3 | mistral  (src/huggingface/ModelLoader.cpp:50-68)
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
    qkv_bias: bool = False   # Qwen2 (src/model/ModelQwen2.h:26-31)
    qk_norm: bool = False    # Qwen3 (src/model/ModelQwen3.h:28-31)
    max_ctx: int = 4096      # rows of the RoPE table / KV-cache capacity used here
    rope_scaling: Optional[RopeScaling] = None

    @property
    def q_dim(self) -> int:
        return self.q_heads * self.head_dim

    @property
    def kv_dim(self) -> int:
        return self.kv_heads * self.head_dim

    @property
    def per_layer_params(self) -> int:
        return (self.q_dim + 2 * self.kv_dim) * self.hidden + self.q_dim * self.hidden + 3 * self.intermediate * self.hidden

    @property
    def weight_params(self) -> int:
        """Matrix parameters one decode step reads (tied lm_head counted once) — SURVEY.md §8a table."""
        return self.layers * self.per_layer_params + self.vocab * self.hidden

    def bytes_per_token(self, ctx: int = 0) -> int:
        """Algorithmic HBM bytes per decoded token: weights once + K/V rows read + new K/V row written."""
        return 2 * self.weight_params + 4 * self.layers * self.kv_dim * ctx + 4 * self.layers * self.kv_dim

    def with_ctx(self, max_ctx: int) -> "ModelSpec":
        return replace(self, max_ctx=max_ctx)


# The BASELINE.json configurations (standard HF configs; table in SURVEY.md §8).
QWEN25_05B = ModelSpec("Qwen2.5-0.5B", "qwen2", 896, 24, 14, 2, 64, 4864, 151936, 1e6, 1e-6, tie=True, qkv_bias=True)
LLAMA32_3B = ModelSpec("Llama-3.2-3B", "llama", 3072, 28, 24, 8, 128, 8192, 128256, 5e5, 1e-5, tie=True,
                       rope_scaling=RopeScaling(32.0, 4.0, 1.0, 8192))
QWEN3_17B = ModelSpec("Qwen3-1.7B", "qwen3", 2048, 28, 16, 8, 128, 6144, 151936, 1e6, 1e-6, tie=True, qk_norm=True)
MISTRAL_7B = ModelSpec("Mistral-7B-v0.3", "mistral", 4096, 32, 32, 8, 128, 14336, 32768, 1e6, 1e-5, tie=False)

# Small shapes for CPU-oracle parity (same code paths: bias / qk-norm / llama3 scaling / GQA group sizes 7,3,2,4).
# Llama, Qwen2 and Mistral derive head_dim = hidden / heads in the reference (src/model/ModelQwen2.h:20,
# ModelLlama.h:37), so hidden = q_heads * head_dim there; Qwen3 carries an explicit head_dim.
TINY_QWEN2 = ModelSpec("tiny-qwen2", "qwen2", 896, 2, 14, 2, 64, 256, 512, 1e6, 1e-6, tie=True, qkv_bias=True,
                       max_ctx=256)
TINY_LLAMA = ModelSpec("tiny-llama", "llama", 768, 2, 6, 2, 128, 512, 640, 5e5, 1e-5, tie=True, max_ctx=256,
                       rope_scaling=RopeScaling(32.0, 4.0, 1.0, 64))
TINY_QWEN3 = ModelSpec("tiny-qwen3", "qwen3", 192, 2, 4, 2, 128, 320, 384, 1e6, 1e-6, tie=True, qk_norm=True,
                       max_ctx=256)
TINY_MISTRAL = ModelSpec("tiny-mistral", "mistral", 1024, 2, 8, 2, 128, 448, 512, 1e6, 1e-5, tie=False, max_ctx=256)

# 8 KV heads: attention heads shard over 2, 4 and 8 tensor-parallel ranks (the other tiny shapes have 2 KV heads)
TINY_TP8 = ModelSpec("tiny-tp8", "mistral", 1024, 2, 8, 8, 128, 512, 1024, 1e6, 1e-5, tie=False, max_ctx=256)

SPECS: Dict[str, ModelSpec] = {s.name: s for s in
                               (QWEN25_05B, LLAMA32_3B, QWEN3_17B, MISTRAL_7B, TINY_QWEN2, TINY_LLAMA, TINY_QWEN3,
                                TINY_MISTRAL, TINY_TP8)}


def synth_weights(spec: ModelSpec, seed: int = 0, device: str = "cpu", std: float = 0.02,
                  device_generator: bool = False) -> Dict[str, torch.Tensor]:
    """Seeded synthetic checkpoint in bf16 with the reference's state names.

    Matrices N(0, std); norm weights 1 + N(0, std) (so the fp32 weight multiply inside RMSNorm is exercised); qkv bias
    N(0, std).  q/k/v and gate/up are stored MERGED ("…qkv_proj.weight", "…gate_up_proj.weight": rows [q|k|v],
    [gate|up]) exactly as the reference's MergedLinear holds them; `split_views()` exposes the HF-named slices.
    Generation is per tensor from a CPU generator so that the same seed gives the same checkpoint on every machine
    (what the oracle-parity tests need).  `device_generator=True` draws on `device` instead (Philox, seconds instead of
    minutes for 7B parameters; identical on every GPU of the same type, which is what tensor-parallel ranks need).
    """
    gen_dev = device if device_generator else "cpu"
    g = torch.Generator(device=gen_dev).manual_seed(seed)

    def mat(*shape):
        return (torch.randn(*shape, generator=g, dtype=torch.float32, device=gen_dev) * std).to(torch.bfloat16).to(device)

    def norm(n):
        return (1.0 + torch.randn(n, generator=g, dtype=torch.float32, device=gen_dev) * std).to(torch.bfloat16).to(device)

    w: Dict[str, torch.Tensor] = {}
    w["model.embed_tokens.weight"] = mat(spec.vocab, spec.hidden)
    for l in range(spec.layers):
        p = f"model.layers.{l}."
        w[p + "input_layernorm.weight"] = norm(spec.hidden)
        w[p + "self_attn.qkv_proj.weight"] = mat(spec.q_dim + 2 * spec.kv_dim, spec.hidden)
        if spec.qkv_bias:
            w[p + "self_attn.qkv_proj.bias"] = mat(spec.q_dim + 2 * spec.kv_dim)
        if spec.qk_norm:
            w[p + "self_attn.q_norm.weight"] = norm(spec.head_dim)
            w[p + "self_attn.k_norm.weight"] = norm(spec.head_dim)
        w[p + "self_attn.o_proj.weight"] = mat(spec.hidden, spec.q_dim)
        w[p + "post_attention_layernorm.weight"] = norm(spec.hidden)
        w[p + "mlp.gate_up_proj.weight"] = mat(2 * spec.intermediate, spec.hidden)
        w[p + "mlp.down_proj.weight"] = mat(spec.hidden, spec.intermediate)
    w["model.norm.weight"] = norm(spec.hidden)
    if not spec.tie:
        w["lm_head.weight"] = mat(spec.vocab, spec.hidden)
    return w


def split_views(spec: ModelSpec, w: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """HF-named state dict (q_proj/k_proj/v_proj, gate_proj/up_proj as dim-0 views of the merged tensors)."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in w.items():
        if k.endswith("self_attn.qkv_proj.weight") or k.endswith("self_attn.qkv_proj.bias"):
            base, kind = k.rsplit("qkv_proj.", 1)
            q, kk, vv = v.split([spec.q_dim, spec.kv_dim, spec.kv_dim], dim=0)
            out[base + "q_proj." + kind], out[base + "k_proj." + kind], out[base + "v_proj." + kind] = q, kk, vv
        elif k.endswith("mlp.gate_up_proj.weight"):
            base = k[: -len("gate_up_proj.weight")]
            gt, up = v.split([spec.intermediate, spec.intermediate], dim=0)
            out[base + "gate_proj.weight"], out[base + "up_proj.weight"] = gt, up
        else:
            out[k] = v
    return out


def rope_table(spec: ModelSpec) -> torch.Tensor:
    """fp32 cos/sin table [max_ctx, head_dim, 2] computed on the host exactly like the reference's ropeInit formulae
    (TT/Operation/OpNNLayerCuda.cuh:359-410): invFreq = 1/powf(θ, 2i/hd) (+ llama3 scaling), angle = pos·invFreq,
    (cosf, sinf) duplicated for both halves.  In a TinyGPT integration the engine is handed the reference's own table
    (RoPE::cache(), TT/Module/Basic.cpp:111) instead."""
    hd, half = spec.head_dim, spec.head_dim // 2
    i = np.arange(half, dtype=np.float32)
    inv = (np.float32(1.0) / np.power(np.float32(spec.rope_theta), (i * np.float32(2.0)) / np.float32(hd),
                                      dtype=np.float32)).astype(np.float32)
    sc = spec.rope_scaling
    if sc is not None:
        orig = np.float32(sc.original_context_length)
        low_wave, high_wave = orig / np.float32(sc.low_freq_factor), orig / np.float32(sc.high_freq_factor)
        wave = np.float32(2.0) * np.float32(np.pi) / inv
        smooth = (orig / wave - np.float32(sc.low_freq_factor)) / (np.float32(sc.high_freq_factor) -
                                                                   np.float32(sc.low_freq_factor))
        scaled = inv / np.float32(sc.factor)
        mid = ((np.float32(1.0) - smooth) * scaled + smooth * inv).astype(np.float32)
        inv = np.where(wave > low_wave, scaled, np.where(wave < high_wave, inv, mid)).astype(np.float32)
    ang = (np.arange(spec.max_ctx, dtype=np.float32)[:, None] * inv[None, :]).astype(np.float32)
    c, s = np.cos(ang, dtype=np.float32), np.sin(ang, dtype=np.float32)
    tab = np.empty((spec.max_ctx, hd, 2), dtype=np.float32)
    tab[:, :half, 0], tab[:, :half, 1], tab[:, half:, 0], tab[:, half:, 1] = c, s, c, s
    return torch.from_numpy(tab)


# ------------------------------------------------------------------------------------------- checkpoint writer
def save_checkpoint(spec: ModelSpec, w: Dict[str, torch.Tensor], out_dir: str) -> None:
    """Write config.json + generation_config.json + model.safetensors in the layout the reference's loader accepts
    (src/huggingface/ModelLoader.cpp:25-87, src/util/SafeTensors.cpp:141-229), so the same synthetic checkpoint can be
    fed to an unmodified TinyGPT build.  Tokenizer files are not written (copy them from the reference's
    assets/tokenizer/<family>)."""
    d = Path(out_dir)
    d.mkdir(parents=True, exist_ok=True)
    cfg = {
        "model_type": spec.model_type, "hidden_size": spec.hidden, "num_hidden_layers": spec.layers,
        "num_attention_heads": spec.q_heads, "num_key_value_heads": spec.kv_heads, "head_dim": spec.head_dim,
        "intermediate_size": spec.intermediate, "vocab_size": spec.vocab, "rope_theta": spec.rope_theta,
        "rms_norm_eps": spec.rms_eps, "tie_word_embeddings": spec.tie, "max_position_embeddings": spec.max_ctx,
        "torch_dtype": "bfloat16", "hidden_act": "silu",
    }
    if spec.rope_scaling is not None:
        sc = spec.rope_scaling
        cfg["rope_scaling"] = {"factor": sc.factor, "high_freq_factor": sc.high_freq_factor,
                               "low_freq_factor": sc.low_freq_factor,
                               "original_max_position_embeddings": sc.original_context_length, "rope_type": "llama3"}
    (d / "config.json").write_text(json.dumps(cfg, indent=1))
    (d / "generation_config.json").write_text(json.dumps({"do_sample": False, "temperature": 0.0, "top_p": 1.0}))
    state = split_views(spec, w)
    header, off, blobs = {}, 0, []
    for name in sorted(state):
        t = state[name].detach().to("cpu").contiguous()
        assert t.dtype == torch.bfloat16
        raw = t.view(torch.int16).numpy().tobytes()
        header[name] = {"dtype": "BF16", "shape": list(t.shape), "data_offsets": [off, off + len(raw)]}
        off += len(raw)
        blobs.append(raw)
    hj = json.dumps(header, separators=(",", ":")).encode()
    hj += b" " * ((8 - len(hj) % 8) % 8)
    with open(d / "model.safetensors", "wb") as f:
        f.write(struct.pack("<Q", len(hj)))
        f.write(hj)
        for b in blobs:
            f.write(b)
