"""Per-launch timeline of one decode token from in-kernel %globaltimer stamps (B200_TRACE=1)."""
import ctypes as C
import os
import sys

os.environ["B200_TRACE"] = "1"
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinygpt_b200 import engine, models  # noqa: E402
from tinygpt_b200._lib import lib  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "Qwen2.5-0.5B"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1          # > 1: the batched token (csrc/gemv_batch*.cu)
spec = models.SPECS[name].with_ctx(256)
w = {k: v.cuda() for k, v in models.synth_weights(spec, seed=0).items()}
eng = engine.DecodeEngine(spec, w)
prompt = torch.randint(0, spec.vocab, (B, 16), generator=torch.Generator().manual_seed(0)).cuda()
eng.reset_cache()
eng.gen_next_token(prompt)
eng.decode(64)
torch.cuda.synchronize()
n = 5 * spec.layers + 1
buf = (C.c_uint64 * (8 * n))()
got = lib().b200_engine_debug_trace(eng._h, buf, n)
names = ["qkv", "attn", "o", "gu", "down"]
t0 = buf[0]
rows = []
for i in range(got):
    a, b, c = buf[8 * i] - t0, buf[8 * i + 1] - t0, buf[8 * i + 2] - t0
    pro = (buf[8 * i + 3] - buf[8 * i + 1]) / 1e3 if buf[8 * i + 3] else 0.0
    if i % 5 == 1 and i < 5 * spec.layers:  # attention: slots 4-5 are %globaltimer stamps
        fine = tuple(((buf[8 * i + j] - buf[8 * i + 1]) / 1e3 if buf[8 * i + j] else 0.0) for j in (4, 5, 6))
    else:                                   # GEMV: slots 4-7 are SM cycles since the dependency resolved
        fine = tuple(buf[8 * i + j] / 1965.0 for j in (4, 5, 6, 7))
    rows.append((names[i % 5] if i < 5 * spec.layers else "head", i // 5, a / 1e3, b / 1e3, c / 1e3, pro, fine))
print(f"{name} (batch {B}): kernel  layer  entry_us  after_wait_us  exit_us   (wait = after_wait-entry, body = exit-after_wait)")
for r in rows[:12] + rows[5 * 10:5 * 10 + 6] + rows[-6:]:
    print(f"{r[0]:5s} {r[1]:3d}  {r[2]:9.2f} {r[3]:9.2f} {r[4]:9.2f}   wait {r[3]-r[2]:6.2f}  body {r[4]-r[3]:6.2f}")
import collections
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0])
prev_exit = None
for r in rows:
    a = agg[r[0]]
    a[0] += 1
    a[1] += r[3] - r[2]
    a[2] += r[4] - r[3]
    a[4] += r[5]
    a[5] += r[6][0]; a[6] += r[6][1]; a[7] += r[6][2]
    a[8] += r[6][3] if len(r[6]) > 3 else 0.0
    if prev_exit is not None:
        a[3] += r[3] - prev_exit  # gap between the previous kernel's exit stamp and this kernel's wait return
    prev_exit = r[4]
print("avg per kernel type: wait-before-dependency, body, gap(prev exit -> my wait return)")
for k, (c, wt, bd, gp, pr, f4, f5, f6, f7) in agg.items():
    if k == "attn":  # slots 3-5 of the attention kernel: q/k/v staged, scores done, softmax done (after the wait)
        print(f"  {k:5s} x{c:3d}: wait {wt/c:6.2f} us  body {bd/c:6.2f} us (q/k/v staged {pr/c:5.2f}, scores {f4/c:5.2f}, "
              f"softmax {f5/c:5.2f}, P.V = body end)  gap {gp/c:6.2f} us")
        continue
    print(f"  {k:5s} x{c:3d}: wait {wt/c:6.2f} us  body {bd/c:6.2f} us (x ready {pr/c:5.2f}; SM clock: 1st stage {f4/c:5.2f}, "
          f"1st block k-loop done {f5/c:5.2f}, stored {f6/c:5.2f}, last block stored {f7/c:5.2f})  gap {gp/c:6.2f} us")
print(f"token span {rows[-1][4]:.1f} us")
