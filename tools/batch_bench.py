"""Batched decode: µs per step and aggregate tokens/s for B = 1, 2, 4, 8 sequences (weights streamed once per step).
    python tools/batch_bench.py [model …]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinygpt_b200 import engine, models  # noqa: E402

for name in (sys.argv[1:] or ["Qwen2.5-0.5B", "Llama-3.2-3B"]):
    spec = models.SPECS[name].with_ctx(256)
    w = models.synth_weights(spec, seed=0, device="cuda", device_generator=True)
    base = None
    for B in (1, 2, 4, 8):
        eng = engine.DecodeEngine(spec, w)
        prompts = torch.randint(0, spec.vocab, (B, 16), generator=torch.Generator().manual_seed(B)).cuda()
        try:
            eng.reset_cache()
            eng.gen_next_token(prompts)
            eng.decode(8)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.decode(96)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 96 * 1e3
            base = base or us
            print(f"{name}: B = {B}: {us:8.1f} us/step  {B / us * 1e6:8.0f} tok/s aggregate  ({us / base:.2f} x the batch-1 step, "
                  f"{B * base / us:.2f} x the throughput of {B} sequential batch-1 steps)", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"{name}: B = {B}: {e}")
        eng.close()
