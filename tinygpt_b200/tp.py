"""Tensor-parallel decode: one process per GPU, Megatron sharding, reductions fused into the GEMV kernels.

Sharding: q/k/v by head and gate/up by FFN column (column-parallel), o_proj/down_proj by input column
(row-parallel, partial hidden vectors summed across ranks), lm_head by vocabulary, embedding + norms replicated.  When
the head counts do not divide by the world size (Qwen2.5-0.5B: 14/2 heads at 4 or 8 GPUs) attention is replicated and
only the FFN and the vocabulary are sharded.  The reference has no tensor-parallel inference (README.md:32); the
single-GPU logits are the oracle for every world size.

torch.distributed is used for plumbing only (exchanging the CUDA-IPC handles of the per-rank exchange windows); the
data path is peer stores over NVLink issued by the GEMV epilogues.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import models
from ._lib import IPC_HANDLE_BYTES, B200Error, ModelDesc, check, lib, require_device
from .engine import DecodeEngine


def can_shard_attention(spec: models.ModelSpec, world: int) -> bool:
    return spec.q_heads % world == 0 and spec.kv_heads % world == 0


def check_shardable(spec: models.ModelSpec, world: int) -> None:
    if spec.intermediate % world or (spec.intermediate // world) % 8:
        raise B200Error(f"{spec.name}: intermediate {spec.intermediate} does not shard over {world} ranks")
    if spec.vocab % world:
        raise B200Error(f"{spec.name}: vocabulary {spec.vocab} does not shard over {world} ranks")


def shard_weights(spec: models.ModelSpec, w: Dict[str, torch.Tensor], rank: int, world: int,
                  shard_attn: Optional[bool] = None) -> Dict[str, torch.Tensor]:
    """This rank's contiguous shards, packed in the engine's merged layout ([q_r|k_r|v_r], [gate_r|up_r])."""
    check_shardable(spec, world)
    if shard_attn is None:
        shard_attn = can_shard_attention(spec, world)
    if shard_attn and not can_shard_attention(spec, world):
        raise B200Error(f"{spec.name}: {spec.q_heads}/{spec.kv_heads} heads do not shard over {world} ranks")
    qd, kvd, I = spec.q_dim, spec.kv_dim, spec.intermediate
    qd_l, kvd_l, I_l, V_l = qd // world, kvd // world, I // world, spec.vocab // world
    out: Dict[str, torch.Tensor] = {}
    for name, t in w.items():
        if name.endswith("self_attn.qkv_proj.weight") or name.endswith("self_attn.qkv_proj.bias"):
            if shard_attn:
                q, k, v = t.split([qd, kvd, kvd], dim=0)
                t = torch.cat([q[rank * qd_l:(rank + 1) * qd_l], k[rank * kvd_l:(rank + 1) * kvd_l],
                               v[rank * kvd_l:(rank + 1) * kvd_l]], dim=0)
        elif name.endswith("self_attn.o_proj.weight"):
            if shard_attn:
                t = t[:, rank * qd_l:(rank + 1) * qd_l]
        elif name.endswith("mlp.gate_up_proj.weight"):
            g, u = t.split([I, I], dim=0)
            t = torch.cat([g[rank * I_l:(rank + 1) * I_l], u[rank * I_l:(rank + 1) * I_l]], dim=0)
        elif name.endswith("mlp.down_proj.weight"):
            t = t[:, rank * I_l:(rank + 1) * I_l]
        elif name == "lm_head.weight":
            t = t[rank * V_l:(rank + 1) * V_l]
        out[name] = t.contiguous()
    if "lm_head.weight" not in out:  # tied: the shard is a row range of the (replicated) embedding — no copy
        out["lm_head.weight"] = out["model.embed_tokens.weight"][rank * V_l:(rank + 1) * V_l]
    return out


class TPDecodeEngine(DecodeEngine):
    """Rank `rank` of a `world`-way tensor-parallel engine.  `weights` are the FULL merged tensors (any device); the
    rank's shards are cut here and moved to the current CUDA device.  Requires an initialised default process group
    (any backend) to exchange the IPC handles."""

    def __init__(self, spec: models.ModelSpec, weights: Dict[str, torch.Tensor], rank: int, world: int,
                 device: torch.device, rope_table: Optional[torch.Tensor] = None, shard_attn: Optional[bool] = None):
        import torch.distributed as dist
        require_device()
        if world < 2 or world > 8:
            raise B200Error("TPDecodeEngine: world must be 2…8")
        if shard_attn is None:
            shard_attn = can_shard_attention(spec, world)
        self.spec, self.device, self.rank, self.world, self.shard_attn = spec, device, rank, world, shard_attn
        self._w = {k: v.to(device) for k, v in shard_weights(spec, weights, rank, world, shard_attn).items()}
        if rope_table is None:  # device-built like the reference's op::ropeInit (see DecodeEngine)
            from . import ops
            with torch.cuda.device(device):
                rope_table = ops.rope_init(spec.head_dim, spec.max_ctx, spec.rope_theta, spec.rope_scaling, device=device)
        self._rope = rope_table.to(device).contiguous()
        self.local_vocab = spec.vocab // world
        desc, table = self._describe(spec, self._w, rank, world, shard_attn)
        # exchange windows: create mine, export, gather everybody's handles, map the peers
        nbytes = lib().b200_tp_window_bytes(C.byref(desc))
        mine = C.c_void_p()
        handle = C.create_string_buffer(IPC_HANDLE_BYTES)
        with torch.cuda.device(device):
            check(lib().b200_tp_window_create(nbytes, C.byref(mine), handle), "b200_tp_window_create")
            handles = [None] * world
            dist.all_gather_object(handles, bytes(handle.raw))
            self._windows = (C.c_void_p * world)()
            self._peer = []
            for r in range(world):
                if r == rank:
                    self._windows[r] = mine
                else:
                    p = C.c_void_p()
                    check(lib().b200_tp_window_open(handles[r], C.byref(p)), "b200_tp_window_open")
                    self._windows[r] = p
                    self._peer.append(p)
            self._mine = mine
            dist.barrier()
            h = C.c_void_p()
            check(lib().b200_engine_create_tp(C.byref(desc), C.byref(table), self._windows, C.byref(h)),
                  "b200_engine_create_tp")
        self._h = h
        dist.barrier()

    def close(self) -> None:
        if getattr(self, "_h", None):
            super().close()
            for p in self._peer:
                lib().b200_tp_window_close(p)
            lib().b200_tp_window_destroy(self._mine)
            self._peer, self._mine = [], None
