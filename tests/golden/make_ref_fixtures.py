"""Generate tests/golden/ref_cpu_ops.npz and ref_cpu_models.npz by running the REFERENCE itself (compiled from
/root/reference into oracle/_ref/libtinygpt_ref.so by oracle/Makefile) on the CPU in fp32.

    make -C oracle && python tests/golden/make_ref_fixtures.py

Only this script needs /root/reference (through the built library); the tests read the committed .npz files.
Inputs are seeded; outputs are the reference's.  The Llama-family forward uses the harness' naive CPU attention shim
(the reference has no CPU flashAttention), everything else is the reference's modules/ops/KV cache/model wiring.
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from tinygpt_b200 import models  # noqa: E402

lib = C.CDLL(str(ROOT / "oracle" / "_ref" / "libtinygpt_ref.so"))
F = C.POINTER(C.c_float)
I64 = C.POINTER(C.c_int64)


def fp(a):
    return a.ctypes.data_as(F)


class Desc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("family", "hidden", "layers", "q_heads", "kv_heads", "head_dim",
                                          "intermediate", "vocab", "max_ctx")] + \
               [("rope_theta", C.c_float), ("rms_eps", C.c_float), ("tie", C.c_int32), ("rs_factor", C.c_float),
                ("rs_high", C.c_float), ("rs_low", C.c_float), ("rs_orig", C.c_int32), ("bf16", C.c_int32)]


lib.ref_model_create.restype = C.c_void_p
lib.ref_model_create.argtypes = [C.POINTER(Desc)]
lib.ref_model_num_states.restype = C.c_int64
lib.ref_model_num_states.argtypes = [C.c_void_p]
lib.ref_model_state_info.restype = C.c_int64
lib.ref_model_state_info.argtypes = [C.c_void_p, C.c_int64, C.c_char_p, C.c_int64]
lib.ref_model_set_state.argtypes = [C.c_void_p, C.c_int64, F]
lib.ref_model_reset.argtypes = [C.c_void_p]
lib.ref_model_forward.argtypes = [C.c_void_p, I64, C.c_int64, F]
lib.ref_model_destroy.argtypes = [C.c_void_p]
lib.ref_rmsnorm_f32.argtypes = [F, F, C.c_float, C.c_int64, C.c_int64, F]
lib.ref_rope_table_f32.argtypes = [C.c_int64, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int64, F]
lib.ref_rope_apply_f32.argtypes = [F] + [C.c_int64] * 4 + [C.c_int, C.c_int64, C.c_int64, C.c_float, C.c_int64, F]
lib.ref_linear_f32.argtypes = [F, F, F] + [C.c_int64] * 4 + [F]
lib.ref_silu_mul_f32.argtypes = [F, C.c_int64, C.c_int64, F]
lib.ref_add_f32.argtypes = [F, F, C.c_int64, F]
lib.ref_argmax_f32.argtypes = [F, C.c_int64, C.c_int64, I64]

rng = np.random.default_rng(1234)
ops = {}

# rmsnorm
x = rng.standard_normal((5, 96)).astype(np.float32)
w = (1 + 0.1 * rng.standard_normal(96)).astype(np.float32)
y = np.empty_like(x)
lib.ref_rmsnorm_f32(fp(x), fp(w), 1e-6, 5, 96, fp(y))
ops.update(rms_x=x, rms_w=w, rms_y=y)

# rope tables: plain and llama3-scaled
for tag, (hd, ctx, theta, sc) in {"plain": (64, 40, 1e6, (0, 0, 0, 0)), "llama3": (128, 48, 5e5, (32.0, 4.0, 1.0, 16))}.items():
    t = np.empty((ctx, hd, 2), np.float32)
    lib.ref_rope_table_f32(hd, ctx, theta, sc[0], sc[1], sc[2], sc[3], fp(t))
    ops[f"rope_table_{tag}"] = t
    ops[f"rope_table_{tag}_args"] = np.array([hd, ctx, theta, *sc], np.float64)

# rope apply, both layouts
xr = rng.standard_normal((2, 3, 4, 64)).astype(np.float32)
for bshd in (0, 1):
    yr = np.empty_like(xr)
    lib.ref_rope_apply_f32(fp(xr), 2, 3, 4, 64, bshd, 64, 40, 1e6, 5, fp(yr))
    ops[f"rope_apply_{'bshd' if bshd else 'bhsd'}"] = yr
ops["rope_apply_x"] = xr

# linear with bias on a 3-D input
xl = rng.standard_normal((1, 3, 40)).astype(np.float32)
Wl = (0.1 * rng.standard_normal((24, 40))).astype(np.float32)
bl = (0.1 * rng.standard_normal(24)).astype(np.float32)
yl = np.empty((1, 3, 24), np.float32)
lib.ref_linear_f32(fp(xl), fp(Wl), fp(bl), 1, 3, 24, 40, fp(yl))
ops.update(lin_x=xl, lin_w=Wl, lin_b=bl, lin_y=yl)

# siluMul, add, argmax (CPU rule)
gu = rng.standard_normal((3, 2 * 20)).astype(np.float32)
ys = np.empty((3, 20), np.float32)
lib.ref_silu_mul_f32(fp(gu), 3, 20, fp(ys))
a, b = rng.standard_normal(50).astype(np.float32), rng.standard_normal(50).astype(np.float32)
ya = np.empty(50, np.float32)
lib.ref_add_f32(fp(a), fp(b), 50, fp(ya))
am_x = rng.standard_normal((4, 33)).astype(np.float32)
am_x[0, 7] = am_x[0, 20] = 9.0  # a tie: the reference CPU path keeps the FIRST maximum (its CUDA path the last)
am = np.empty(4, np.int64)
lib.ref_argmax_f32(fp(am_x), 4, 33, am.ctypes.data_as(I64))
ops.update(silu_gu=gu, silu_y=ys, add_a=a, add_b=b, add_y=ya, argmax_x=am_x, argmax_cpu=am)
np.savez_compressed(ROOT / "tests" / "golden" / "ref_cpu_ops.npz", **ops)

# whole models: tiny specs, fp32 CPU, prompt + teacher-forced decode steps
FAMILY = {"llama": 0, "qwen2": 1, "qwen3": 2, "mistral": 3}
out = {}
for spec in (models.TINY_QWEN2, models.TINY_LLAMA, models.TINY_QWEN3, models.TINY_MISTRAL):
    sc = spec.rope_scaling
    d = Desc(FAMILY[spec.model_type], spec.hidden, spec.layers, spec.q_heads, spec.kv_heads, spec.head_dim,
             spec.intermediate, spec.vocab, spec.max_ctx, spec.rope_theta, spec.rms_eps, int(spec.tie),
             sc.factor if sc else 0.0, sc.high_freq_factor if sc else 0.0, sc.low_freq_factor if sc else 0.0,
             sc.original_context_length if sc else 0, 0)
    h = lib.ref_model_create(C.byref(d))
    state = models.split_views(spec, models.synth_weights(spec, seed=0))
    n = lib.ref_model_num_states(h)
    name = C.create_string_buffer(256)
    seen = []
    for i in range(n):
        cnt = lib.ref_model_state_info(h, i, name, 256)
        key = name.value.decode()
        seen.append(key)
        if key.endswith(".rope"):
            continue  # the reference's own table
        src = state.get(key)
        if src is None and key == "lm_head.weight" and spec.tie:
            src = state["model.embed_tokens.weight"]
        assert src is not None, f"reference state {key} has no synthetic tensor"
        arr = np.ascontiguousarray(src.float().numpy().reshape(-1))
        assert arr.size == cnt, (key, arr.size, cnt)
        lib.ref_model_set_state(h, i, fp(arr))
    g = torch.Generator().manual_seed(5)
    prompt = torch.randint(0, spec.vocab, (7,), generator=g, dtype=torch.int64).numpy()
    forced = torch.randint(0, spec.vocab, (5,), generator=g, dtype=torch.int64).numpy()
    logits = np.empty((6, spec.vocab), np.float32)
    lib.ref_model_reset(h)
    lib.ref_model_forward(h, prompt.ctypes.data_as(I64), 7, fp(logits[0]))
    for t in range(5):
        tok = np.array([forced[t]], np.int64)
        lib.ref_model_forward(h, tok.ctypes.data_as(I64), 1, fp(logits[t + 1]))
    out[f"{spec.name}.prompt"], out[f"{spec.name}.forced"], out[f"{spec.name}.logits"] = prompt, forced, logits
    out[f"{spec.name}.state_names"] = np.array(seen)
    print(spec.name, "states", n, "logits std", float(logits.std()))
np.savez_compressed(ROOT / "tests" / "golden" / "ref_cpu_models.npz", **out)
print("wrote fixtures")


# ---------------------------------------------------------------------------------------------------- sampler
# The reference's own Sampler::sample (src/engine/Sampler.cpp, compiled unmodified into the library) with the CPU
# `multinomial` op hooked through the registry (oracle/ref_harness.cpp ref_sampler_f32): records the probabilities the
# sampler hands to multinomial and draws by inverse CDF from a given uniform number.
lib.ref_sampler_f32.restype = C.c_int64
lib.ref_sampler_f32.argtypes = [F, C.c_int64, C.c_float, C.c_int64, C.c_float, C.c_float, C.c_float, F]
srng = np.random.default_rng(4321)
SAMPLER_CFGS = [(0.8, 0, 1.0, 0.0), (1.0, 50, 1.0, 0.0), (0.7, 0, 0.9, 0.0), (1.3, 0, 1.0, 0.05), (0.6, 40, 0.95, 0.02),
                (2.0, 5, 0.5, 0.0), (1.0, 1, 1.0, 0.0), (0.9, 0, 0.0001, 0.0), (1.0, 100000, 0.999, 0.5),
                (0.0, 0, 1.0, 0.0)]
SAMPLER_U = [0.0001, 0.37, 0.62, 0.9999]
samp = {"cfgs": np.array(SAMPLER_CFGS, dtype=np.float64), "u": np.array(SAMPLER_U, dtype=np.float32)}
for V in (97, 2048):
    logits = (srng.standard_normal(V) * 3).astype(np.float32)
    samp[f"logits_{V}"] = logits
    probs = np.zeros((len(SAMPLER_CFGS), V), dtype=np.float32)
    picks = np.zeros((len(SAMPLER_CFGS), len(SAMPLER_U)), dtype=np.int64)
    for ci, (T, k, tp_, mp) in enumerate(SAMPLER_CFGS):
        for ui, u in enumerate(SAMPLER_U):
            out = np.zeros(V, dtype=np.float32)
            picks[ci, ui] = lib.ref_sampler_f32(fp(logits), V, T, k, tp_, mp, u, fp(out))
            probs[ci] = out
    samp[f"probs_{V}"] = probs
    samp[f"picks_{V}"] = picks
np.savez_compressed(ROOT / "tests" / "golden" / "ref_cpu_sampler.npz", **samp)
print("wrote ref_cpu_sampler.npz")
import os
os._exit(0)  # skip the reference's static teardown (allocator asserts on destruction order)
