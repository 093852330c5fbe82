"""BASELINE config 4: prefill S tokens (tcgen05 GEMM path) + decode N tokens, timed with CUDA events.
    python tools/prefill_bench.py [model] [S] [N]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinygpt_b200 import engine, models  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "Qwen3-1.7B"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
N = int(sys.argv[3]) if len(sys.argv) > 3 else 256
spec = models.SPECS[name].with_ctx(S + N + 16)
w = models.synth_weights(spec, seed=0, device="cuda", device_generator=True)
eng = engine.DecodeEngine(spec, w)
prompt = torch.randint(0, spec.vocab, (1, S), generator=torch.Generator().manual_seed(0)).cuda()
for rep in range(3):
    eng.reset_cache()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    eng.gen_next_token(prompt)
    e1.record()
    toks = eng.decode(N)
    e2.record()
    torch.cuda.synchronize()
    t_pre, t_dec = e0.elapsed_time(e1), e1.elapsed_time(e2)
    gemm_flop = 2.0 * S * spec.layers * spec.per_layer_params + 2.0 * spec.vocab * spec.hidden
    attn_flop = 2.0 * spec.layers * S * S * spec.q_dim  # causal half of 4·S²·qDim
    ctx_mid = S + N // 2
    print(f"{name}: prefill {S} tokens {t_pre:.2f} ms = {(gemm_flop + attn_flop) / t_pre / 1e9:.1f} TFLOP/s "
          f"({S / t_pre * 1e3:.0f} prompt tok/s); decode {N} tokens {t_dec / N * 1e3:.1f} us/token = "
          f"{N / t_dec * 1e3:.0f} tok/s, {spec.bytes_per_token(ctx_mid) / (t_dec / N) / 1e6:.0f} GB/s at ctx {S}→{S + N}")
