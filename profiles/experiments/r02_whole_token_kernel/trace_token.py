"""Per-op, per-CTA timeline of one decode token inside the whole-token persistent kernel (mega.cu), from in-kernel
%globaltimer stamps (B200_TRACE=1): op entered → activation vector staged → op done, for every CTA."""
import ctypes as C
import os
import sys

os.environ["B200_TRACE"] = "1"
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinygpt_b200 import engine, models  # noqa: E402
from tinygpt_b200._lib import lib  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "Qwen2.5-0.5B"
spec = models.SPECS[name].with_ctx(256)
w = models.synth_weights(spec, seed=0, device="cuda", device_generator=True)
eng = engine.DecodeEngine(spec, w)
prompt = torch.randint(0, spec.vocab, (1, 16), generator=torch.Generator().manual_seed(0)).cuda()
eng.reset_cache()
eng.gen_next_token(prompt)
eng.decode(64)
torch.cuda.synchronize()
grid = torch.cuda.get_device_properties(0).multi_processor_count
rows = 5 * spec.layers + 2
words = grid * rows * 4
buf = (C.c_uint64 * words)()
got = lib().b200_engine_debug_trace_token(eng._h, buf, words)
if got <= 0:
    raise SystemExit("no trace (engine not running the whole-token kernel?)")
t = np.frombuffer(buf, dtype=np.uint64).reshape(grid, rows, 4).astype(np.float64) / 1e3   # µs
t0 = t[:, rows - 1, 0].min()
tok = t[:, rows - 1, 1].max() - t0
n_ops = rows - 1
names = ["qkv", "attn", "o", "gu", "down"]
print(f"{name}: token {tok:.1f} us ({n_ops} ops, {grid} CTAs)")
print("per op kind, averaged over layers (µs): CTAs taking part | op span = last done − first enter | stage x: median, max | "
      "rows: median, max | done spread (last − first CTA done)")
agg = {}
for j in range(n_ops):
    kind = names[j % 5] if j < 5 * spec.layers else "head"
    ent, xs, done = t[:, j, 0], t[:, j, 1], t[:, j, 2]
    part = xs > 0
    if not part.any():
        continue
    a = agg.setdefault(kind, [])
    a.append((part.sum(), done[part].max() - ent[part].min(), np.median((xs - ent)[part]), (xs - ent)[part].max(),
              np.median((done - xs)[part]), (done - xs)[part].max(), done[part].max() - done[part].min(),
              done[part].max() - t0))
for k, v in agg.items():
    m = np.mean(np.array(v)[:, :7], axis=0)
    print(f"  {k:5s} x{len(v):3d}: {m[0]:5.0f} CTAs | span {m[1]:6.2f} | x {m[2]:5.2f} {m[3]:5.2f} | rows {m[4]:5.2f} {m[5]:5.2f} | "
          f"spread {m[6]:5.2f}")
# critical path: time between consecutive ops' "last CTA done"
last_done = [max(t[:, j, 2].max(), 0) - t0 for j in range(n_ops)]
per = {}
prev = 0.0
for j in range(n_ops):
    kind = names[j % 5] if j < 5 * spec.layers else "head"
    per.setdefault(kind, []).append(last_done[j] - prev)
    prev = last_done[j]
print("critical path: Δ between successive ops' last-CTA-done, mean per kind:",
      {k: round(float(np.mean(v)), 2) for k, v in per.items()})
np.save(os.environ.get("B200_TRACE_OUT", "/tmp/trace_token.npy"), t)
