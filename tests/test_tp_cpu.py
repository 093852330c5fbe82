"""Host-side logic of the tensor-parallel path on the CPU: two `gloo` ranks shard a synthetic checkpoint with
tinygpt_b200.tp.shard_weights, run the oracle's arithmetic on their shards with the exchange the engine performs
(fp32 partial hidden vectors summed in rank order, residual + roundings applied after the sum; vocabulary-sharded argmax
merged with the last-index tie rule) and must reproduce the unsharded oracle."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import orc, to_oracle_cfg
from tinygpt_b200 import models, tp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _sharded_forward(spec, w_local, ids, table, rank, world, shard_attn):
    """One prefill through the TP dataflow with oracle primitives; returns (local logits of the last position)."""
    cfg = to_oracle_cfg(spec)
    Hq_l = spec.q_heads // world if shard_attn else spec.q_heads
    Hkv_l = spec.kv_heads // world if shard_attn else spec.kv_heads
    qd_l, kvd_l = Hq_l * spec.head_dim, Hkv_l * spec.head_dim

    def allreduce_fp32(partial):  # the engine: every rank sums the same fp32 partials in rank order
        parts = [torch.zeros_like(partial) for _ in range(world)]
        dist.all_gather(parts, partial.contiguous())
        acc = torch.zeros_like(partial)
        for p in parts:
            acc = acc + p
        return acc

    x = orc.embedding(w_local["model.embed_tokens.weight"], ids)
    B, S, _ = x.shape
    for l in range(spec.layers):
        p = f"model.layers.{l}."
        h = orc.rms_norm(x, w_local[p + "input_layernorm.weight"], spec.rms_eps)
        qkv = orc.linear(h, w_local[p + "self_attn.qkv_proj.weight"], w_local.get(p + "self_attn.qkv_proj.bias"))
        q = qkv[..., :qd_l].reshape(B, S, Hq_l, spec.head_dim)
        k = qkv[..., qd_l:qd_l + kvd_l].reshape(B, S, Hkv_l, spec.head_dim)
        v = qkv[..., qd_l + kvd_l:].reshape(B, S, Hkv_l, spec.head_dim)
        if spec.qk_norm:
            q = orc.rms_norm(q, w_local[p + "self_attn.q_norm.weight"], spec.rms_eps)
            k = orc.rms_norm(k, w_local[p + "self_attn.k_norm.weight"], spec.rms_eps)
        q, k = orc.rope_apply(q, table, 0), orc.rope_apply(k, table, 0)
        o = orc.flash_attention(q, k, v, True).reshape(B, S, qd_l)
        acc = o.float() @ w_local[p + "self_attn.o_proj.weight"].float().t()           # fp32 partial, NOT rounded
        if shard_attn:
            acc = allreduce_fp32(acc)
        x = orc.add(x, orc.rnd(acc, "bf16"))
        h = orc.rms_norm(x, w_local[p + "post_attention_layernorm.weight"], spec.rms_eps)
        m = orc.silu_mul(orc.linear(h, w_local[p + "mlp.gate_up_proj.weight"], None))
        acc = allreduce_fp32(m.float() @ w_local[p + "mlp.down_proj.weight"].float().t())
        x = orc.add(x, orc.rnd(acc, "bf16"))
    x = orc.rms_norm(x, w_local["model.norm.weight"], spec.rms_eps)
    return orc.linear(x, w_local["lm_head.weight"], None)[0, -1]


def _worker(rank, world, port, spec_name, shard_attn, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        spec = models.SPECS[spec_name]
        w = models.synth_weights(spec, seed=0)
        table = models.rope_table(spec)
        ids = torch.arange(3, 12).view(1, -1)
        local = _sharded_forward(spec, tp.shard_weights(spec, w, rank, world, shard_attn), ids, table, rank, world,
                                 shard_attn)
        # vocabulary-sharded argmax with the reference tie rule: (max, GLOBAL index) per rank, highest index wins ties
        V_l = spec.vocab // world
        val, idx = float(local.max()), int(orc.argmax_last(local.view(1, -1))) + rank * V_l
        cands = [None] * world
        dist.all_gather_object(cands, (val, idx))
        best = max(cands, key=lambda c: (c[0], c[1]))
        shards = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(shards, local)
        if rank == 0:
            q.put((torch.cat(shards), best[1]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("spec_name,shard_attn", [("tiny-mistral", True), ("tiny-qwen3", True), ("tiny-qwen2", True),
                                                  ("tiny-qwen2", False), ("tiny-llama", True)])
def test_tp2_dataflow_matches_unsharded_oracle(spec_name, shard_attn):
    world = 2
    ctx = mp.get_context("spawn")
    result = None
    for attempt in range(2):  # the probed port can be taken between the probe and the rendezvous: retry once
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, spec_name, shard_attn, q)) for r in range(world)]
        for p in procs:
            p.start()
        try:
            result = q.get(timeout=240)
        except Exception:
            result = None
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.terminate()
        if result is not None and all(p.exitcode == 0 for p in procs):
            break
        result = None
    assert result is not None, "two gloo ranks did not complete (twice)"
    logits_tp, tok_tp = result
    spec = models.SPECS[spec_name]
    w = models.synth_weights(spec, seed=0)
    ids = torch.arange(3, 12).view(1, -1)
    want = orc.forward(to_oracle_cfg(spec), w, ids, orc.KVCache(), models.rope_table(spec), "bf16")[0, -1]
    # same rounding points; only the fp32 summation order of the two row-parallel GEMVs differs
    assert float((logits_tp - want).abs().mean()) < 2e-3
    assert float((logits_tp - want).abs().max()) < 3e-2
    assert tok_tp == int(orc.argmax_last(logits_tp.view(1, -1)))


def test_shard_weights_partition_is_exact():
    for spec in (models.TINY_MISTRAL, models.TINY_QWEN2, models.TINY_QWEN3):
        w = models.synth_weights(spec, seed=1)
        for world in (2,):
            sh = [tp.shard_weights(spec, w, r, world) for r in range(world)]
            p = "model.layers.1."
            I_l = spec.intermediate // world
            gu = w[p + "mlp.gate_up_proj.weight"]
            assert torch.equal(torch.cat([s[p + "mlp.gate_up_proj.weight"][:I_l] for s in sh]), gu[:spec.intermediate])
            assert torch.equal(torch.cat([s[p + "mlp.gate_up_proj.weight"][I_l:] for s in sh]), gu[spec.intermediate:])
            assert torch.equal(torch.cat([s[p + "mlp.down_proj.weight"] for s in sh], dim=1), w[p + "mlp.down_proj.weight"])
            assert torch.equal(torch.cat([s[p + "self_attn.o_proj.weight"] for s in sh], dim=1),
                               w[p + "self_attn.o_proj.weight"])
            head = w.get("lm_head.weight", w["model.embed_tokens.weight"])
            assert torch.equal(torch.cat([s["lm_head.weight"] for s in sh]), head)
            qd_l = spec.q_dim // world
            assert torch.equal(torch.cat([s[p + "self_attn.qkv_proj.weight"][:qd_l] for s in sh]),
                               w[p + "self_attn.qkv_proj.weight"][:spec.q_dim])


def test_unshardable_configs_are_rejected():
    from tinygpt_b200._lib import B200Error
    assert not tp.can_shard_attention(models.QWEN25_05B, 4)      # 14 / 2 heads: only 2-way
    assert tp.can_shard_attention(models.QWEN25_05B, 2)
    assert all(tp.can_shard_attention(models.MISTRAL_7B, n) for n in (2, 4, 8))
    with pytest.raises(B200Error):
        tp.shard_weights(models.TINY_QWEN2, models.synth_weights(models.TINY_QWEN2), 0, 4, shard_attn=True)
