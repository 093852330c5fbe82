"""Weight-loader fast path: HF checkpoint directory → the engine's merged, per-rank device layout.

What the reference does  [ref: src/huggingface/ModelLoader.cpp:25-87, src/huggingface/ModelConfig.cpp:73-122,
src/util/SafeTensors.cpp:141-229 (single file), :231-300 (index of shards), src/layer/Linear.h:64-79 (merged views)]:
build the module tree, mmap the file, copy every tensor whole into its (merged-view) destination, then run a
`model().to(dtype)` pass; there is no way to load a tensor-parallel slice, and every RoPE module owns its own table.

Here, for the decode engine:
  * `load_model_config`  — config.json → ModelSpec with the reference's defaults and derivations (head_dim =
    hidden/heads for llama/qwen2/mistral, explicit for qwen3; qkv bias for qwen2; q/k-norm for qwen3; llama3 rope
    scaling only for llama).
  * `SafeTensorsFile`    — header parse + mmap; tensors are zero-copy numpy views of the file.
  * `load_checkpoint`    — allocates each destination tensor ONCE in the engine's merged layout ([q|k|v], [gate|up]) and
    copies only the row / column ranges THIS rank owns straight from the mapping (row ranges are contiguous in the
    file; column ranges of o_proj / down_proj are strided reads), through a reusable pinned staging buffer when the
    destination is a CUDA device.  No full-size intermediate, no dtype conversion pass (the file dtype must already be
    bf16, which is also what the reference requires: SafeTensors.cpp:196-201), one RoPE table for the whole model.
The result equals tp.shard_weights(spec, full_weights, rank, world) bit for bit (tests/test_loader.py).

Error behaviour mirrors the reference's strict load: a missing tensor, a shape or dtype mismatch fail the load
(LoaderError); unexpected tensors are ignored with a warning list; `lm_head.weight` may be absent when the embeddings
are tied (and is ignored when present).
"""
from __future__ import annotations

import json
import mmap
import struct
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import models
from ._lib import B200Error


class LoaderError(B200Error):
    pass


_SUPPORTED = ("llama", "qwen2", "qwen3", "mistral")


def load_model_config(path: str | Path, max_ctx: Optional[int] = None) -> models.ModelSpec:
    """config.json → ModelSpec  [ref: src/huggingface/ModelConfig.cpp:45-122; src/model/ModelLlama.h:21-42,
    ModelQwen2.h:20-34, ModelQwen3.h:25-31, ModelMistral.h:25-29]."""
    p = Path(path)
    if p.is_dir():
        p = p / "config.json"
    try:
        doc = json.loads(p.read_text())
    except Exception as e:
        raise LoaderError(f"Failed to load model config: {p}: {e}")
    mt = doc.get("model_type", "")
    if mt not in _SUPPORTED:
        raise LoaderError(f"Unsupported model_type: {mt!r} (the decode engine covers {', '.join(_SUPPORTED)})")
    need = ("hidden_size", "num_hidden_layers", "num_attention_heads", "num_key_value_heads", "intermediate_size",
            "vocab_size")
    missing = [k for k in need if int(doc.get(k, -1)) <= 0]
    if missing:
        raise LoaderError(f"config.json: missing or non-positive {missing}")
    dt = doc.get("torch_dtype", doc.get("dtype", ""))
    if dt not in ("bfloat16",):
        raise LoaderError(f"config.json: torch_dtype {dt!r}; the B200 decode path is bf16")
    H, heads = int(doc["hidden_size"]), int(doc["num_attention_heads"])
    # qwen3 carries an explicit head_dim; the other families derive it (ModelLlama.h:37, ModelQwen2.h:25, ModelMistral.h:25)
    hd = int(doc.get("head_dim", -1)) if mt == "qwen3" else H // heads
    if hd <= 0:
        raise LoaderError("config.json: head_dim missing for qwen3")
    scaling = None
    if mt == "llama" and isinstance(doc.get("rope_scaling"), dict):
        # The reference applies the llama3 formula to EVERY llama config that carries a rope_scaling object, whatever its
        # rope_type (ModelConfig.cpp:79-88 parses the four numbers, ModelLlama.h:21-25,42 always passes them on).  Other
        # rope types (linear, dynamic, yarn) would silently get llama3 arithmetic there; here they are refused.
        rs = doc["rope_scaling"]
        rt = rs.get("rope_type", rs.get("type", "llama3"))
        if rt != "llama3":
            raise LoaderError(f"config.json: rope_scaling.rope_type {rt!r}: only llama3 scaling is built "
                              "(the reference would apply the llama3 formula to it)")
        scaling = models.RopeScaling(float(rs.get("factor", 1.0)), float(rs.get("high_freq_factor", 1.0)),
                                     float(rs.get("low_freq_factor", 1.0)),
                                     int(rs.get("original_max_position_embeddings", -1)))
    theta_default = 1.0 if mt == "llama" else 10000.0
    ctx = int(max_ctx if max_ctx is not None else doc.get("max_position_embeddings", 4096))
    return models.ModelSpec(
        name=str(doc.get("_name_or_path") or p.parent.name), model_type=mt, hidden=H,
        layers=int(doc["num_hidden_layers"]), q_heads=heads, kv_heads=int(doc["num_key_value_heads"]), head_dim=hd,
        intermediate=int(doc["intermediate_size"]), vocab=int(doc["vocab_size"]),
        rope_theta=float(doc.get("rope_theta", theta_default)), rms_eps=float(doc.get("rms_norm_eps", 1e-5)),
        tie=bool(doc.get("tie_word_embeddings", False)), qkv_bias=(mt == "qwen2"), qk_norm=(mt == "qwen3"),
        max_ctx=ctx, rope_scaling=scaling)


class SafeTensorsFile:
    """One .safetensors file: 8-byte little-endian header length, JSON header, raw tensor bytes; mmapped read-only."""

    def __init__(self, path: str | Path):
        self.path = Path(path)
        try:
            self._f = open(self.path, "rb")
            self._mm = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ)
        except Exception as e:
            raise LoaderError(f"Error mapFileForRead: {self.path}: {e}")
        if len(self._mm) < 8:
            raise LoaderError(f"{self.path}: not a safetensors file")
        (hlen,) = struct.unpack("<Q", self._mm[:8])
        if hlen <= 0 or 8 + hlen > len(self._mm):
            raise LoaderError(f"{self.path}: bad header length {hlen}")
        try:
            self.header: Dict[str, dict] = json.loads(bytes(self._mm[8:8 + hlen]).decode("utf-8"))
        except Exception as e:
            raise LoaderError(f"{self.path}: header is not JSON: {e}")
        self.header.pop("__metadata__", None)
        self._base = 8 + hlen

    def keys(self) -> List[str]:
        return list(self.header)

    def tensor_u16(self, name: str, shape: Tuple[int, ...]) -> np.ndarray:
        """Zero-copy uint16 view [shape] of a BF16 tensor; shape, dtype and byte count are checked like the reference's
        strict load (SafeTensors.cpp:186-212)."""
        info = self.header[name]
        if info.get("dtype") != "BF16":
            raise LoaderError(f"dtype not equal for tensor: {name} (file has {info.get('dtype')}, engine wants BF16)")
        if tuple(info.get("shape", ())) != tuple(shape):
            raise LoaderError(f"shape not equal for tensor: {name} (file {info.get('shape')}, model {list(shape)})")
        start, end = info["data_offsets"]
        n = int(np.prod(shape)) if len(shape) else 1
        if end - start != 2 * n or self._base + end > len(self._mm):
            raise LoaderError(f"size not equal for tensor: {name}")
        return np.frombuffer(self._mm, dtype=np.uint16, count=n, offset=self._base + start).reshape(shape)

    def close(self) -> None:
        # numpy views keep the mapping alive; closing is best effort
        try:
            self._mm.close()
        except BufferError:
            pass
        self._f.close()


class _Checkpoint:
    """All tensors of a checkpoint directory (single file or an index of shards)."""

    def __init__(self, model_dir: str | Path):
        d = Path(model_dir)
        self.files: Dict[str, SafeTensorsFile] = {}
        self.where: Dict[str, str] = {}
        single, index = d / "model.safetensors", d / "model.safetensors.index.json"
        if single.exists():
            f = SafeTensorsFile(single)
            self.files[single.name] = f
            self.where = {k: single.name for k in f.keys()}
        elif index.exists():
            try:
                wm = json.loads(index.read_text())["weight_map"]
            except Exception as e:
                raise LoaderError(f"Error open index file: {index}: {e}")
            for name, fname in wm.items():
                if fname not in self.files:
                    self.files[fname] = SafeTensorsFile(d / fname)
                if name not in self.files[fname].header:
                    raise LoaderError(f"{index}: {name} is not in {fname}")
                self.where[name] = fname
        else:
            raise LoaderError(f"Load model failed: neither {single} nor {index} exists")

    def get(self, name: str, shape: Tuple[int, ...]) -> np.ndarray:
        if name not in self.where:
            raise LoaderError(f"Missing key: {name}")
        return self.files[self.where[name]].tensor_u16(name, shape)

    def close(self) -> None:
        for f in self.files.values():
            f.close()


class _Copier:
    """uint16 numpy slice → bf16 destination slice; CUDA destinations go through one reusable pinned staging buffer so
    the H2D copies are asynchronous DMA from page-locked memory instead of a pageable memcpy per tensor.
    `staging_bytes` on a CPU destination forces the same chunked path through an ordinary buffer (tests)."""

    def __init__(self, device: torch.device, staging_bytes: Optional[int] = None):
        self.device = device
        self.cuda = device.type == "cuda"
        self.staging = None
        if self.cuda:
            self.staging = torch.empty((staging_bytes or (64 << 20)) // 2, dtype=torch.int16).pin_memory()
        elif staging_bytes:
            self.staging = torch.empty(staging_bytes // 2, dtype=torch.int16)
        self.bytes = 0

    def _sync(self) -> None:
        if self.cuda:
            torch.cuda.current_stream(self.device).synchronize()

    def copy(self, dst: torch.Tensor, src: np.ndarray) -> None:
        """dst: contiguous bf16 view with src's shape (a dim-0 range of the destination tensor)."""
        assert tuple(dst.shape) == tuple(src.shape) and dst.is_contiguous()
        self.bytes += src.size * 2
        d16 = dst.view(torch.int16).view(-1)
        if self.staging is None:
            d16.copy_(torch.from_numpy(np.array(src, dtype=np.uint16, copy=True).view(np.int16).reshape(-1)))
            return
        rows = src.shape[0] if src.ndim > 1 else 1
        row_elems = src.size // max(rows, 1)
        s2 = src.reshape(rows, row_elems) if src.ndim != 2 else src
        cap = self.staging.numel()
        if row_elems > cap:                       # a single row larger than the buffer: walk it in column pieces
            for r in range(rows):
                for c0 in range(0, row_elems, cap):
                    c1 = min(row_elems, c0 + cap)
                    self._sync()                  # the previous piece has left the staging buffer
                    np.copyto(self.staging[:c1 - c0].numpy().view(np.uint16), s2[r, c0:c1])
                    d16[r * row_elems + c0:r * row_elems + c1].copy_(self.staging[:c1 - c0], non_blocking=True)
            self._sync()
            return
        cap_rows = cap // row_elems
        for r0 in range(0, rows, cap_rows):
            r1 = min(rows, r0 + cap_rows)
            n = (r1 - r0) * row_elems
            self._sync()                          # the previous chunk has left the staging buffer
            stage = self.staging[:n].numpy().view(np.uint16).reshape(r1 - r0, row_elems)
            np.copyto(stage, s2[r0:r1])           # the only host-side pass: strided (column slice) or not
            d16[r0 * row_elems:r0 * row_elems + n].copy_(self.staging[:n], non_blocking=True)
        self._sync()


def load_checkpoint(model_dir: str | Path, device: str | torch.device = "cuda", rank: int = 0, world: int = 1,
                    shard_attn: Optional[bool] = None, max_ctx: Optional[int] = None,
                    strict_unexpected: bool = False,
                    staging_bytes: Optional[int] = None) -> Tuple[models.ModelSpec, Dict[str, torch.Tensor], dict]:
    """→ (spec, weights in the engine's merged layout holding rank `rank`'s shards on `device`, report).

    The dict plugs straight into DecodeEngine (world == 1) or is what TPDecodeEngine would cut for itself
    (tp.shard_weights); report = {'bytes': copied, 'unexpected': [...], 'files': n}."""
    from . import tp  # sharding rules live there
    spec = load_model_config(model_dir, max_ctx)
    dev = torch.device(device)
    if world < 1 or rank < 0 or rank >= world:
        raise LoaderError(f"bad rank {rank} / world {world}")
    if world > 1:
        try:
            tp.check_shardable(spec, world)
        except B200Error as e:
            raise LoaderError(str(e))
        if shard_attn is None:
            shard_attn = tp.can_shard_attention(spec, world)
        if shard_attn and not tp.can_shard_attention(spec, world):
            raise LoaderError(f"{spec.name}: {spec.q_heads}/{spec.kv_heads} heads do not shard over {world} ranks")
    else:
        shard_attn = False
    ck = _Checkpoint(model_dir)
    cp = _Copier(dev, staging_bytes)
    H, qd, kvd, I, V = spec.hidden, spec.q_dim, spec.kv_dim, spec.intermediate, spec.vocab
    qd_l, kvd_l = (qd // world, kvd // world) if shard_attn else (qd, kvd)
    a_rank = rank if shard_attn else 0                  # attention slices (replicated ⇒ everything from offset 0)
    I_l, V_l = I // world, V // world
    used = set()
    w: Dict[str, torch.Tensor] = {}

    def new(*shape):
        return torch.empty(*shape, dtype=torch.bfloat16, device=dev)

    def src(name, shape):
        used.add(name)
        return ck.get(name, shape)

    def whole(name, shape):
        t = new(*shape)
        cp.copy(t, src(name, shape))
        return t

    try:
        w["model.embed_tokens.weight"] = whole("model.embed_tokens.weight", (V, H))
        for l in range(spec.layers):
            p = f"model.layers.{l}."
            w[p + "input_layernorm.weight"] = whole(p + "input_layernorm.weight", (H,))
            # merged [q_r | k_r | v_r]: three row ranges of three file tensors into one allocation
            qkv = new(qd_l + 2 * kvd_l, H)
            cp.copy(qkv[:qd_l], src(p + "self_attn.q_proj.weight", (qd, H))[a_rank * qd_l:(a_rank + 1) * qd_l])
            cp.copy(qkv[qd_l:qd_l + kvd_l],
                    src(p + "self_attn.k_proj.weight", (kvd, H))[a_rank * kvd_l:(a_rank + 1) * kvd_l])
            cp.copy(qkv[qd_l + kvd_l:],
                    src(p + "self_attn.v_proj.weight", (kvd, H))[a_rank * kvd_l:(a_rank + 1) * kvd_l])
            w[p + "self_attn.qkv_proj.weight"] = qkv
            if spec.qkv_bias:
                b = new(qd_l + 2 * kvd_l)
                cp.copy(b[:qd_l], src(p + "self_attn.q_proj.bias", (qd,))[a_rank * qd_l:(a_rank + 1) * qd_l])
                cp.copy(b[qd_l:qd_l + kvd_l], src(p + "self_attn.k_proj.bias", (kvd,))[a_rank * kvd_l:(a_rank + 1) * kvd_l])
                cp.copy(b[qd_l + kvd_l:], src(p + "self_attn.v_proj.bias", (kvd,))[a_rank * kvd_l:(a_rank + 1) * kvd_l])
                w[p + "self_attn.qkv_proj.bias"] = b
            if spec.qk_norm:
                w[p + "self_attn.q_norm.weight"] = whole(p + "self_attn.q_norm.weight", (spec.head_dim,))
                w[p + "self_attn.k_norm.weight"] = whole(p + "self_attn.k_norm.weight", (spec.head_dim,))
            o = new(H, qd_l)                                            # row-parallel: a column range (strided read)
            cp.copy(o, src(p + "self_attn.o_proj.weight", (H, qd))[:, a_rank * qd_l:(a_rank + 1) * qd_l])
            w[p + "self_attn.o_proj.weight"] = o
            w[p + "post_attention_layernorm.weight"] = whole(p + "post_attention_layernorm.weight", (H,))
            gu = new(2 * I_l, H)                                        # merged [gate_r | up_r]
            cp.copy(gu[:I_l], src(p + "mlp.gate_proj.weight", (I, H))[rank * I_l:(rank + 1) * I_l])
            cp.copy(gu[I_l:], src(p + "mlp.up_proj.weight", (I, H))[rank * I_l:(rank + 1) * I_l])
            w[p + "mlp.gate_up_proj.weight"] = gu
            dn = new(H, I_l)
            cp.copy(dn, src(p + "mlp.down_proj.weight", (H, I))[:, rank * I_l:(rank + 1) * I_l])
            w[p + "mlp.down_proj.weight"] = dn
        w["model.norm.weight"] = whole("model.norm.weight", (H,))
        if spec.tie:
            # tied: the head is (a row range of) the embedding — no second copy [ref: src/model/GPTModel.h:43-48]
            if world > 1:
                w["lm_head.weight"] = w["model.embed_tokens.weight"][rank * V_l:(rank + 1) * V_l]
            used.add("lm_head.weight")
        else:
            head = new(V_l, H)
            cp.copy(head, src("lm_head.weight", (V, H))[rank * V_l:(rank + 1) * V_l])
            w["lm_head.weight"] = head
        unexpected = sorted(k for k in ck.where if k not in used)
        if unexpected and strict_unexpected:
            raise LoaderError(f"Unexpected key: {unexpected[0]} (+{len(unexpected) - 1} more)")
        return spec, w, {"bytes": cp.bytes, "unexpected": unexpected, "files": len(ck.files)}
    finally:
        ck.close()
