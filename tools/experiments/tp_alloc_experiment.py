"""Why is the one-GPU engine slower inside a torchrun job than alone (Mistral-7B: 3 780 vs 2 810 µs/token)?
Times the SAME single-GPU decode (a) before the NCCL process group exists, (b) after it, same allocations, (c) after it,
fresh allocations, (d) after the tensor-parallel IPC windows were created and mapped.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/experiments/tp_alloc_experiment.py [model]
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tinygpt_b200 import engine, models, tp  # noqa: E402


def timed(eng, prompt, n=128):
    eng.reset_cache()
    eng.gen_next_token(prompt)
    eng.decode(8)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        eng.seek(prompt.shape[1])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.decode(n)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return best


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "Mistral-7B-v0.3"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    spec = models.SPECS[name].with_ctx(160)
    prompt = torch.randint(0, spec.vocab, (1, 16), generator=torch.Generator().manual_seed(0)).to(dev)
    w = models.synth_weights(spec, seed=0, device=dev, device_generator=True)
    eng = engine.DecodeEngine(spec, w)
    rows = [("before init_process_group", timed(eng, prompt))]
    dist.init_process_group("nccl", device_id=dev)
    x = torch.ones(8, device=dev)
    dist.all_reduce(x)
    torch.cuda.synchronize()
    rows.append(("after NCCL init, same allocations", timed(eng, prompt)))
    eng.close()
    del eng, w
    torch.cuda.empty_cache()
    w = models.synth_weights(spec, seed=0, device=dev, device_generator=True)
    eng = engine.DecodeEngine(spec, w)
    rows.append(("after NCCL init, fresh allocations", timed(eng, prompt)))
    dist.barrier()
    small = models.SPECS["tiny-tp8"]
    teng = tp.TPDecodeEngine(small, models.synth_weights(small, seed=0), rank, world, dev)
    rows.append(("after TP IPC windows mapped (cudaIpcOpenMemHandle), same allocations", timed(eng, prompt)))
    eng.close()
    del eng, w
    torch.cuda.empty_cache()
    w = models.synth_weights(spec, seed=0, device=dev, device_generator=True)
    eng = engine.DecodeEngine(spec, w)
    rows.append(("after TP IPC windows mapped, fresh allocations", timed(eng, prompt)))
    # other ranks idle from here: does a peer's traffic matter?
    dist.barrier()
    if rank == 0:
        rows.append(("rank 0 alone, peers idle at a store wait", timed(eng, prompt)))
        dist.distributed_c10d._get_default_store().set("done", "1")
    else:
        import datetime
        dist.distributed_c10d._get_default_store().wait(["done"], datetime.timedelta(seconds=300))
    if rank == 0:
        for k, v in rows:
            print(f"{name}: {v:8.1f} us/token  {k}", flush=True)
    teng.close()
    eng.close()
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
