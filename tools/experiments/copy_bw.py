"""Copy bandwidth of every visible GPU (1 GiB bf16, read + write bytes, best of 10 — the method behind
MEASURED_PEAKS.json hbm_gbs) and clocks, so that numbers taken on different boxes of the pool can be compared."""
import subprocess

import torch

for i in range(torch.cuda.device_count()):
    torch.cuda.set_device(i)
    a = torch.empty(1 << 29, dtype=torch.bfloat16, device="cuda")
    b = torch.empty_like(a)
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        b.copy_(a)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"GPU {i}: copy {2 * a.numel() * 2 / best / 1e6:.0f} GB/s", flush=True)
    del a, b
print(subprocess.run(["nvidia-smi", "--query-gpu=index,name,clocks.sm,clocks.mem,clocks.max.mem,power.limit,ecc.mode.current",
                      "--format=csv"], capture_output=True, text=True).stdout)
