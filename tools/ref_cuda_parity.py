"""Engine and oracle against the REFERENCE'S OWN CUDA PATH on the same GPU (north_star: "logits matching the reference
CUDA path within 1e-3 (bf16) / bit-exact for argmax token ids on the same prompts").

oracle/_ref/ref_cuda_decode is the unmodified reference (TinyTorch CUDA ops + cuBLAS + TinyFA + its KV-cache manager +
its argmax) compiled from /root/reference by `make -C oracle cuda` in the build container; it travels to the GPU box as
a binary.  For every model this script
  1. writes the seeded synthetic checkpoint in the reference loader's layout (models.save_checkpoint),
  2. runs our engine (free-running greedy decode, then logits teacher-forced on its own tokens),
  3. runs the reference CUDA path teacher-forced on the SAME tokens, and once free-running,
  4. runs the CPU oracle teacher-forced on the same tokens (small models),
and reports |engine − reference|, |oracle − reference| (this pins the oracle's bf16 rounding points, the one item that
could not be pinned without a GPU), greedy-id agreement, and the reference's own decode speed next to ours.

    python tools/ref_cuda_parity.py [--models tiny-qwen2,tiny-llama,tiny-qwen3,tiny-mistral,Qwen2.5-0.5B] [--new 24]
                                    [--json gpurun_out/ref_cuda_parity.json]
"""
from __future__ import annotations

import argparse
import json
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from tinygpt_b200 import engine, models  # noqa: E402

REF_BIN = ROOT / "oracle" / "_ref" / "ref_cuda_decode"
FAMILY = {"llama": "llama", "qwen2": "qwen2", "qwen3": "qwen3", "mistral": "mistral"}


def run_reference(spec, ckpt_dir, prompt, n_new, forced=None, time_steps=0, timeout=1200, b200="off", batched=False,
                  dump_rope=None, batch=1):
    """→ (tokens [n_new] int64, logits [n_new, V] float32, timing dict or None) from the reference CUDA binary.
    b200 = "ops" / "engine": the same reference program with our kernels behind its op registry (boundary B) / our
    engine behind GPTModel::model() (boundary A), through integration/tinytorch_b200_adapter.h."""
    d = Path(ckpt_dir)
    (d / "ids.bin").write_bytes(np.asarray(prompt, dtype=np.int64).tobytes())
    cmd = [str(REF_BIN), "--ckpt", str(d), "--family", FAMILY[spec.model_type], "--dims",
           ",".join(str(int(v)) for v in (spec.hidden, spec.layers, spec.q_heads, spec.kv_heads, spec.head_dim,
                                          spec.intermediate, spec.vocab, spec.max_ctx)),
           "--theta", repr(float(spec.rope_theta)), "--eps", repr(float(spec.rms_eps)), "--tie", str(int(spec.tie)),
           "--ids", str(d / "ids.bin"), "--new", str(n_new), "--out", str(d / "ref_out.bin"), "--b200", b200,
           "--qkv-bias", str(int(spec.qkv_bias)), "--qk-norm", str(int(spec.qk_norm))]
    if spec.rope_scaling is not None:
        sc = spec.rope_scaling
        cmd += ["--rope-scaling", f"{sc.factor},{sc.high_freq_factor},{sc.low_freq_factor},{sc.original_context_length}"]
    if forced is not None:
        (d / "forced.bin").write_bytes(np.asarray(forced, dtype=np.int64).tobytes())
        cmd += ["--forced", str(d / "forced.bin")]
    if time_steps:
        cmd += ["--time-steps", str(time_steps)]
    if batched:
        cmd += ["--batched", "1"]
    if dump_rope:
        cmd += ["--dump-rope", str(dump_rope)]
    if batch > 1:   # prompt: [B][S] ids, forced: [N][B]; returns tokens [N, B], logits [N, B, V]
        cmd += ["--batch", str(batch)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"ref_cuda_decode failed ({r.returncode}): {r.stderr[-800:]}")
    raw = (d / "ref_out.bin").read_bytes()
    toks = np.frombuffer(raw[: 8 * n_new * batch], dtype=np.int64).copy()
    logits = np.frombuffer(raw[8 * n_new * batch:], dtype=np.float32).copy()
    if batch > 1:
        toks, logits = toks.reshape(n_new, batch), logits.reshape(n_new, batch, spec.vocab)
    else:
        logits = logits.reshape(n_new, spec.vocab)
    timing = None
    for line in r.stdout.splitlines():
        if line.startswith("{"):
            timing = json.loads(line)
    return torch.from_numpy(toks), torch.from_numpy(logits), timing


def run_engine(spec, w, prompt, n_new):
    dev = "cuda"
    eng = engine.DecodeEngine(spec, {k: v.to(dev) for k, v in w.items()})
    p = torch.tensor(prompt, dtype=torch.int64).view(1, -1).to(dev)
    eng.reset_cache()
    first = eng.gen_next_token(p)
    toks = torch.cat([first.view(-1), eng.decode(n_new - 1)]).cpu()
    eng.reset_cache()
    logits = [eng.forward(p)[0, -1].float().cpu()]
    for i in range(n_new - 1):
        logits.append(eng.forward(toks[i].view(1, 1).to(dev))[0, -1].float().cpu())
    # decode speed, device-resident loop (what bench.py times)
    eng.reset_cache()
    eng.gen_next_token(p)
    n_time = min(128, spec.max_ctx - len(prompt) - 2)
    eng.decode(4)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.seek(len(prompt))
    e0.record()
    eng.decode(n_time)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n_time
    eng.close()
    return toks, torch.stack(logits), us


def compare(name, a, b):
    d = (a - b).abs()
    return {"pair": name, "mean_abs": float(d.mean()), "max_abs": float(d.max()),
            "frac_bit_identical": float((a == b).float().mean()), "frac_within_1e-3": float((d <= 1e-3).float().mean())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--models", default="tiny-qwen2,tiny-llama,tiny-qwen3,tiny-mistral,Qwen2.5-0.5B,Qwen3-1.7B,"
                                        "Llama-3.2-3B,Mistral-7B-v0.3")
    ap.add_argument("--oracle-max-params", type=float, default=4e9,
                    help="run the CPU oracle (fp32 copies of the weights) only for models below this size")
    ap.add_argument("--new", type=int, default=24)
    ap.add_argument("--prompt", type=int, default=16)
    ap.add_argument("--json", default="")
    ap.add_argument("--prefill", default="", help="MODEL:TOKENS, e.g. Qwen3-1.7B:2048 — time the prefill of both paths")
    args = ap.parse_args()
    if not REF_BIN.exists():
        raise SystemExit(f"{REF_BIN} is missing: build it in the container that has /root/reference (make -C oracle cuda)")
    from helpers import orc, to_oracle_cfg
    report = []
    for name in args.models.split(","):
        spec = models.SPECS[name]
        if spec.max_ctx > 512:
            spec = spec.with_ctx(256)
        big = spec.hidden * spec.layers > 10000
        w = models.synth_weights(spec, seed=0, device="cuda" if big else "cpu", device_generator=big)
        prompt = torch.randint(0, spec.vocab, (args.prompt,), generator=torch.Generator().manual_seed(0)).tolist()
        with tempfile.TemporaryDirectory() as td:
            models.save_checkpoint(spec, w, td)
            toks, logits, us = run_engine(spec, w, prompt, args.new)
            ref_toks_forced, ref_logits, _ = run_reference(spec, td, prompt, args.new, forced=toks.tolist())
            ref_toks_free, _, timing = run_reference(spec, td, prompt, args.new, time_steps=128 if big else 32)
            # the reference against itself: the same positions through its batched path (one forward over prompt +
            # forced tokens: cuBLAS GEMM m > 1, causal TinyFA) — the noise floor of its own CUDA arithmetic
            _, ref_logits_batched, _ = run_reference(spec, td, prompt, args.new, forced=toks.tolist(), batched=True,
                                                     dump_rope=Path(td) / "rope.bin")
            raw = (Path(td) / "rope.bin").read_bytes()
            shp = np.frombuffer(raw[:24], dtype=np.int64)
            ref_rope = torch.from_numpy(np.frombuffer(raw[24:], dtype=np.float32).reshape(*shp).copy())
        row = {"model": name, "engine_vs_reference_cuda": compare("engine-ref", logits, ref_logits),
               "reference_cuda_decode_vs_its_own_batched_path": compare("ref-ref", ref_logits_batched, ref_logits)}
        from tinygpt_b200 import ops as b200ops
        rows = min(int(shp[0]), spec.max_ctx)
        dev_tab = b200ops.rope_init(spec.head_dim, rows, spec.rope_theta, spec.rope_scaling).cpu()
        host_tab = models.rope_table(spec)[:rows]
        row["rope_table"] = {"reference_rows": int(shp[0]), "compared_rows": rows,
                             "b200_rope_init_bit_identical": bool(torch.equal(dev_tab, ref_rope[:rows])),
                             "b200_rope_init_max_abs": float((dev_tab - ref_rope[:rows]).abs().max()),
                             "host_numpy_table_frac_identical": float((host_tab == ref_rope[:rows]).float().mean()),
                             "host_numpy_table_max_abs": float((host_tab - ref_rope[:rows]).abs().max())}
        # greedy ids: the reference's argmax on ITS logits, forced on our tokens, step by step
        top2 = torch.topk(ref_logits, 2, dim=-1).values
        margin = (top2[:, 0] - top2[:, 1])
        # one bf16 ulp at the magnitude of the winning logit (8 significand bits): a margin ≤ 1 ulp is a near-tie — the
        # two candidates are adjacent bf16 values (or equal), and any summation-order difference can swap them
        ulp = torch.pow(2.0, torch.floor(torch.log2(top2[:, 0].abs().clamp_min(1e-30))) - 7)
        near = margin <= ulp
        agree = (ref_toks_forced == toks)
        row["greedy_ids"] = {"equal_steps": int(agree.sum()), "steps": int(len(toks)),
                             "near_tie_steps_margin_le_1ulp": int(near.sum()),
                             "exact_tie_steps": int((margin == 0).sum()),
                             "different_steps_that_are_near_ties": int((near & ~agree).sum()),
                             "different_steps_with_decisive_margin": int((~near & ~agree).sum()),
                             "margins_in_ulp_where_different": [round(float(m / u), 3) for m, u in
                                                                zip(margin[~agree], ulp[~agree])],
                             "min_margin_where_different": float(margin[~agree].min()) if (~agree).any() else None,
                             "median_margin_in_ulp": float((margin / ulp).median()),
                             "free_running_identical_prefix": int((torch.cumprod((ref_toks_free == toks).long(), 0)).sum())}
        n_params = sum(v.numel() for v in w.values())
        if n_params <= args.oracle_max_params:
            wf = {k: v.float().cpu() for k, v in w.items()}
            _, logits_orc = orc.generate_greedy(to_oracle_cfg(spec), wf, torch.tensor(prompt), args.new,
                                                ref_rope[:rows].contiguous(), "bf16", forced=toks)  # the reference's table
            row["oracle_vs_reference_cuda"] = compare("oracle-ref", logits_orc, ref_logits)
            row["engine_vs_oracle"] = compare("engine-oracle", logits, logits_orc)
            del wf
        row["decode_us_per_token"] = {"ours": us, "reference_cuda": timing["us_per_token"] if timing else None}
        # the drop-in boundary with the real reference around it: same program, our code behind its seams
        with tempfile.TemporaryDirectory() as td:
            models.save_checkpoint(spec, w, td)
            for mode in ("ops", "engine"):
                try:
                    t_m, l_m, tim = run_reference(spec, td, prompt, args.new, forced=toks.tolist(), b200=mode,
                                                  time_steps=128 if big else 32)
                    row[f"reference_with_b200_{mode}"] = {
                        "vs_plain_reference": compare(f"{mode}-ref", l_m, ref_logits),
                        "vs_our_python_engine": compare(f"{mode}-engine", l_m, logits),
                        "greedy_ids_equal_to_ours": int((t_m == toks).sum()), "steps": int(len(toks)),
                        "us_per_token_through_the_reference_loop": tim["us_per_token"] if tim else None}
                except Exception as e:  # noqa: BLE001
                    row[f"reference_with_b200_{mode}"] = {"failed": str(e)[:400]}
        report.append(row)
        print(json.dumps(row), flush=True)
    # config 4 of BASELINE.json: prefill of a long prompt, the reference's own path vs ours (timing only)
    if args.prefill:
        name, plen = args.prefill.split(":")
        plen = int(plen)
        spec = models.SPECS[name].with_ctx(plen + 64)
        w = models.synth_weights(spec, seed=0, device="cuda", device_generator=True)
        prompt = torch.randint(0, spec.vocab, (plen,), generator=torch.Generator().manual_seed(0)).tolist()
        row = {"model": name, "prefill_tokens": plen}
        eng = engine.DecodeEngine(spec, w)
        p = torch.tensor(prompt, dtype=torch.int64).view(1, -1).cuda()
        for _ in range(2):
            eng.reset_cache()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.gen_next_token(p)
            e1.record()
            torch.cuda.synchronize()
        row["ours_prefill_ms"] = e0.elapsed_time(e1)
        eng.close()
        with tempfile.TemporaryDirectory() as td:
            models.save_checkpoint(spec, w, td)
            try:
                _, _, tim = run_reference(spec, td, prompt, 2, time_steps=16)
                row["reference_cuda_prefill_ms"] = tim.get("prefill_ms") if tim else None
                row["reference_cuda_decode_us_per_token_at_this_ctx"] = tim.get("us_per_token") if tim else None
            except Exception as e:  # noqa: BLE001
                row["reference_cuda"] = {"failed": str(e)[:400]}
        report.append(row)
        print(json.dumps(row), flush=True)
    if args.json:
        Path(args.json).parent.mkdir(parents=True, exist_ok=True)
        Path(args.json).write_text(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
