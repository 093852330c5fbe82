"""What would north_star's literal design cost?  It names "a single NCCL all-reduce on the hidden vector per layer"; the
engine instead fuses the reduction into the GEMV kernels (fp32 partials pushed into peer windows as {value, tag} words,
DESIGN.md §6).  This measures the NCCL side of that comparison on the same box: latency of an in-graph
`ncclAllReduce` of one hidden vector (bf16 H and fp32 H elements, H of the four BASELINE models), back to back on one
stream inside a CUDA graph like a decode token would issue them, so that the per-token cost of 2·L all-reduces can be
set against the measured exchange cost of the fused path (tools/trace_tp.py).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/nccl_allreduce_bench.py
"""
import os

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    models = {"Qwen2.5-0.5B": (896, 24), "Qwen3-1.7B": (2048, 28), "Llama-3.2-3B": (3072, 28), "Mistral-7B-v0.3": (4096, 32)}
    rows = []
    for name, (H, L) in models.items():
        for dtype in (torch.bfloat16, torch.float32):
            x = torch.ones(H, dtype=dtype, device=dev)
            n_ar = 2 * L                       # two reductions per layer for exact parity
            for _ in range(5):
                dist.all_reduce(x)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                with torch.cuda.graph(g, stream=stream):
                    for _ in range(n_ar):
                        dist.all_reduce(x)
            for _ in range(3):
                g.replay()
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            e0.record()
            for _ in range(reps):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            us_per_ar = float(t.item()) * 1e3 / (reps * n_ar)
            rows.append((name, H, str(dtype).split(".")[-1], n_ar, us_per_ar, us_per_ar * n_ar))
    if rank == 0:
        print(f"NCCL all-reduce of one hidden vector, in-graph, back to back, {world} GPUs (max over ranks)")
        print("| model | H | dtype | all-reduces / token | us each | us / token |")
        print("|---|---|---|---|---|---|")
        for r in rows:
            print(f"| {r[0]} | {r[1]} | {r[2]} | {r[3]} | {r[4]:.2f} | {r[5]:.0f} |", flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)   # tearing down a process group that holds captured NCCL graphs can block; everything is printed


if __name__ == "__main__":
    main()
