"""Decode a few tokens between cudaProfilerStart/Stop (for ncu --profile-from-start off) and print a quick timing.

    python tools/profile_decode.py [model] [n_profiled_tokens] [ctx_tokens_before]
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinygpt_b200 import engine, models  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "Qwen2.5-0.5B"
n_prof = int(sys.argv[2]) if len(sys.argv) > 2 else 2
pre = int(sys.argv[3]) if len(sys.argv) > 3 else 80
spec = models.SPECS[name].with_ctx(max(256, pre + 200))
w = models.synth_weights(spec, seed=0, device="cuda", device_generator=True)
eng = engine.DecodeEngine(spec, w)
prompt = torch.randint(0, spec.vocab, (1, 16), generator=torch.Generator().manual_seed(0)).cuda()
eng.reset_cache()
eng.gen_next_token(prompt)
eng.decode(pre - 16)
torch.cuda.synchronize()
# quick timing (not under the profiler's range)
for rep in range(3):
    eng.seek(pre)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.decode(64)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 64
    print(f"{name}: {ms * 1e3:.1f} us/token  {1e3 / ms:.0f} tok/s  "
          f"{spec.bytes_per_token(pre + 32) / ms / 1e6:.0f} GB/s  launches/token {eng.launches_per_token}")
eng.seek(pre)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
eng.decode(n_prof)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
