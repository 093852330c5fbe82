"""Per-launch timeline of one TP decode token (rank 0) from in-kernel %globaltimer stamps.  usage: trace_tp.py model world"""
import ctypes as C
import os
import sys

os.environ["B200_TRACE"] = "1"
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def worker(rank, world, name):
    from tinygpt_b200 import models, tp
    from tinygpt_b200._lib import lib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    spec = models.SPECS[name].with_ctx(256)
    eng = tp.TPDecodeEngine(spec, models.synth_weights(spec, seed=0, device=dev, device_generator=True), rank, world, dev)
    prompt = torch.randint(0, spec.vocab, (1, 16), generator=torch.Generator().manual_seed(0)).to(dev)
    eng.reset_cache()
    eng.gen_next_token(prompt)
    eng.decode(64)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.decode(32); e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"{name} tp{world}: {e0.elapsed_time(e1) / 32 * 1e3:.1f} us/token")
        n = 5 * spec.layers + 1
        buf = (C.c_uint64 * (8 * n))()
        got = lib().b200_engine_debug_trace(eng._h, buf, n)
        names = ["qkv", "attn", "o", "gu", "down"]
        import collections
        agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
        prev = None
        for i in range(got):
            a, b, c, d = (buf[8 * i + j] for j in range(4))
            k = names[i % 5] if i < 5 * spec.layers else "head"
            g = agg[k]
            g[0] += 1; g[1] += (b - a) / 1e3; g[2] += (c - b) / 1e3
            g[4] += ((d - b) / 1e3) if d else 0.0
            if prev is not None:
                g[3] += (b - prev) / 1e3
            prev = c
        for k, (c, wt, bd, gp, pr) in agg.items():
            print(f"  {k:5s} x{c:3d}: wait {wt/c:6.2f}  body {bd/c:6.2f} (prologue {pr/c:5.2f})  gap {gp/c:6.2f} us")
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "Qwen2.5-0.5B"
    world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    mp.spawn(worker, args=(world, name), nprocs=world, join=True)
