"""Pin the oracle (oracle/decode_oracle.py) before trusting it — CPU only.

1. against the reference's OWN known-answer vectors (third_party/TinyTorch/test/*.cpp, transcribed with file:line into
   tests/golden/reference_vectors.json), tolerance 1e-3 abs like the reference's test harness;
2. against outputs of the reference ITSELF, compiled from /root/reference by oracle/Makefile and run on the CPU in fp32
   by tests/golden/make_ref_fixtures.py (committed as tests/golden/ref_cpu_*.npz): per-op and whole-model logits of
   the four Llama-family wirings, prefill + teacher-forced decode steps.
"""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from helpers import orc, to_oracle_cfg
from tinygpt_b200 import models

G = Path(__file__).resolve().parent / "golden"
VEC = json.loads((G / "reference_vectors.json").read_text())


def near(a, b, tol=1e-3):
    a, b = torch.as_tensor(a, dtype=torch.float32).reshape(-1), torch.as_tensor(b, dtype=torch.float32).reshape(-1)
    assert a.shape == b.shape
    assert float((a - b).abs().max()) <= tol, float((a - b).abs().max())


# ---------------------------------------------------------------------------------- reference known-answer tests
def test_golden_rmsnorm():
    v = VEC["func_rmsNorm"]
    near(orc.rms_norm(torch.tensor(v["x"]), torch.tensor(v["w"]), v["eps"], "fp32"), v["y"])


def test_golden_silu():
    v = VEC["func_silu"]
    near(orc.silu(torch.tensor(v["x"]), "fp32"), v["y"])
    gu = torch.cat([torch.tensor(v["x"]), torch.ones(4)]).view(1, 8)
    near(orc.silu_mul(gu, "fp32"), v["y"])


def test_golden_linear():
    v = VEC["func_linear"]
    y = orc.linear(torch.tensor(v["x"]), torch.tensor(v["w"]), torch.tensor(v["b"]), "fp32", three_d=False)
    assert abs(float(y.sum()) - v["sum"]) < 1e-4


def test_golden_attention_semantics():
    """sdpAttention golden (BHSD) pins the attention semantics incl. the causal mask; the flash restatement must agree."""
    v = VEC["func_sdpAttention"]
    shp = v["shape_bhsd"]
    q, k, vv = (torch.tensor(v[n]).view(shp).transpose(1, 2).contiguous() for n in ("q", "k", "v"))  # → BSHD
    for causal, key in ((False, "y"), (True, "y_causal")):
        want = torch.tensor(v[key]).view(shp).transpose(1, 2)
        near(orc.naive_attention(q, k, vv, causal), want)
        near(orc.flash_attention(q, k, vv, causal, "fp32"), want)


def test_golden_rope():
    v = VEC["module_rope"]
    near(orc.rope_table(4, 3, 1000.0), v["table_hd4_ctx3_theta1000"], 1e-4)
    sc = v["scaling"]
    near(orc.rope_table(4, 3, 10000.0, orc.RopeScaling(sc["factor"], sc["high_freq_factor"], sc["low_freq_factor"],
                                                       sc["original_context_length"])),
         v["table_scaled_hd4_ctx3_theta10000"], 1e-4)
    # the reference test applies the SECOND (llama3-scaled, θ = 10000) module it built (test_module.cpp:70-83)
    scaled = orc.rope_table(4, 3, 10000.0, orc.RopeScaling(sc["factor"], sc["high_freq_factor"],
                                                          sc["low_freq_factor"], sc["original_context_length"]))
    x = torch.tensor(v["apply_x"]).view(v["apply_shape_bhsd"])
    y = orc.rope_apply(x, scaled, 0, "BHSD", "fp32")
    near(y, v["apply_y"])
    # the same data seen as BSHD gives the same rotation per (head, position)
    y2 = orc.rope_apply(x.transpose(1, 2).contiguous(), scaled, 0, "BSHD", "fp32")
    near(y2.transpose(1, 2), v["apply_y"])


def test_argmax_tie_rule_and_cpu_rule_differ():
    ops = np.load(G / "ref_cpu_ops.npz")
    x = torch.from_numpy(ops["argmax_x"])
    cpu_rule = torch.from_numpy(ops["argmax_cpu"])
    assert torch.equal(torch.argmax(x, dim=-1), cpu_rule), "the reference CPU path keeps the first maximum"
    got = orc.argmax_last(x)
    assert int(got[0]) == 20 and int(cpu_rule[0]) == 7, "CUDA rule: the highest index among equal maxima"
    assert torch.equal(got[1:], cpu_rule[1:])


# -------------------------------------------------------------------------- outputs of the reference itself (fp32)
def test_ref_run_ops():
    o = np.load(G / "ref_cpu_ops.npz")
    t = lambda k: torch.from_numpy(o[k])
    near(orc.rms_norm(t("rms_x"), t("rms_w"), 1e-6, "fp32"), t("rms_y"), 1e-5)
    for tag in ("plain", "llama3"):
        hd, ctx, theta, f, hi, lo, orig = o[f"rope_table_{tag}_args"].tolist()
        sc = None if f == 0 else orc.RopeScaling(f, hi, lo, int(orig))
        near(orc.rope_table(int(hd), int(ctx), theta, sc), t(f"rope_table_{tag}"), 2e-6)
    tab = orc.rope_table(64, 40, 1e6)
    near(orc.rope_apply(t("rope_apply_x"), tab, 5, "BSHD", "fp32"), t("rope_apply_bshd"), 1e-5)
    near(orc.rope_apply(t("rope_apply_x"), tab, 5, "BHSD", "fp32"), t("rope_apply_bhsd"), 1e-5)
    near(orc.linear(t("lin_x"), t("lin_w"), t("lin_b"), "fp32"), t("lin_y"), 1e-5)
    near(orc.silu_mul(t("silu_gu"), "fp32"), t("silu_y"), 1e-6)
    near(orc.add(t("add_a"), t("add_b"), "fp32"), t("add_y"), 0.0)


@pytest.mark.parametrize("spec", [models.TINY_QWEN2, models.TINY_LLAMA, models.TINY_QWEN3, models.TINY_MISTRAL],
                         ids=lambda s: s.name)
def test_ref_run_whole_model_fp32(spec):
    """Same seeded checkpoint, same ids → the reference's fp32 CPU logits (prefill + 5 teacher-forced decode steps
    through its own KV cache) vs the oracle in fp32 mode.  Pins wiring, merged-weight layout, bias, q/k-norm, RoPE
    offset handling, llama3 scaling, GQA mapping, causal-on-prefill / non-causal-on-decode, tied lm_head."""
    m = np.load(G / "ref_cpu_models.npz")
    prompt = torch.from_numpy(m[f"{spec.name}.prompt"])
    forced = torch.from_numpy(m[f"{spec.name}.forced"])
    want = torch.from_numpy(m[f"{spec.name}.logits"])
    w = {k: v.float() for k, v in models.synth_weights(spec, seed=0).items()}
    cfg, table = to_oracle_cfg(spec), models.rope_table(spec)
    cache = orc.KVCache()
    got = [orc.forward(cfg, w, prompt.view(1, -1), cache, table, "fp32")[0, -1]]
    for t in forced:
        got.append(orc.forward(cfg, w, t.view(1, 1), cache, table, "fp32")[0, -1])
    got = torch.stack(got)
    err = float((got - want).abs().max())
    assert err <= 2e-4, f"{spec.name}: oracle(fp32) vs reference CPU fp32 logits differ by {err}"
    assert torch.equal(got.argmax(-1), want.argmax(-1))


def test_state_names_match_reference():
    """The HF state names our checkpoints use are exactly the reference's Module::namedStates() (minus its rope tables)."""
    m = np.load(G / "ref_cpu_models.npz")
    for spec in (models.TINY_QWEN2, models.TINY_LLAMA, models.TINY_QWEN3, models.TINY_MISTRAL):
        ref_names = {n for n in m[f"{spec.name}.state_names"].tolist() if not n.endswith(".rope")}
        ours = set(models.split_views(spec, models.synth_weights(spec, seed=0)).keys())
        if spec.tie:
            ours.add("lm_head.weight")  # aliases embed_tokens in the reference (GPTModel.h:39-41)
        assert ref_names == ours, (spec.name, sorted(ref_names ^ ours))


def test_bf16_rounding_points_are_exercised():
    """bf16 mode differs from fp32 mode (the roundings are really applied) but stays close to it."""
    spec = models.TINY_QWEN2
    w = models.synth_weights(spec, seed=0)
    cfg, table = to_oracle_cfg(spec), models.rope_table(spec)
    ids = torch.arange(1, 8).view(1, -1)
    a = orc.forward(cfg, w, ids, orc.KVCache(), table, "bf16")
    b = orc.forward(cfg, {k: v.float() for k, v in w.items()}, ids, orc.KVCache(), table, "fp32")
    d = float((a - b).abs().max())
    assert 0 < d < 0.1
    assert torch.equal(a, a.to(torch.bfloat16).float()), "bf16-mode logits are bf16 values"
