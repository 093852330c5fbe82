"""tcgen05 prefill GEMM alone, at the Linear shapes of BASELINE config 4 (Qwen3-1.7B, 2 048 prompt tokens): TFLOP/s of
the one-tile-per-CTA kernel and of the persistent 128×256 kernel (B200_GEMM=persistent) against the measured cuBLAS
bf16 peak in MEASURED_PEAKS.json.  CUDA events on the current stream, operands cycled so that B never sits in L2."""
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tinygpt_b200 import ops  # noqa: E402

peak = 1661.0
try:
    peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["bf16_tflops"])
except Exception:
    pass
H, I, QKV = 2048, 6144, 4096
shapes = [("qkv", QKV, H), ("o", H, 2048), ("gate_up", 2 * I, H), ("down", H, I)]
for M in (512, 2048):
    for name, N, K in shapes:
        a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        ws = [(torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16) for _ in range(8)]  # 8 × ≥ 8 MB ≫ reuse in L2
        row = [f"M={M:5d} {name:8s} N={N:6d} K={K:5d}"]
        for mode in ("tile", "persistent"):
            os.environ["B200_GEMM"] = mode
            for w in ws[:3]:
                ops.gemm(a, w)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record()
            for _ in range(reps):
                for w in ws:
                    ops.gemm(a, w)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / (reps * len(ws))
            tf = 2.0 * M * N * K / ms / 1e9
            row.append(f"{mode}: {ms * 1e3:7.1f} us {tf:7.1f} TFLOP/s ({tf / peak:.2f} of measured peak)")
        os.environ.pop("B200_GEMM", None)
        print("  ".join(row))
