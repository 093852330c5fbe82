#!/usr/bin/env python
"""bench.py — decode tokens/s (bf16, batch 1) of the B200 decode engine, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--model Qwen2.5-0.5B] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is ONE 128-token greedy decode of one sequence (BASELINE.json configs[1]: Qwen2.5-0.5B bf16, batch 1,
128-token decode) continuing a 16-token prompt: rewind to position 16, run 128 engine steps on the device.
`value` = decoded tokens / second with everything resident in HBM (CUDA events on the launching stream).
`e2e`   = the same through the public API with HOST buffers: DecodeEngine.generate_sync(prompt ids on the host) —
          H2D of the prompt, cache reset, 16-token prefill, 127 decode steps, D2H of the 128 ids, all inside the clock.
Synthetic seeded weights of the real shape (no checkpoints offline); weights (0.99 GB) ≫ L2 (126 MB), so every token
re-streams them from HBM — no explicit L2 flush is needed and none is done.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

PROMPT_LEN = 16
NEW_TOKENS = 128


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (one streaming `nvidia-smi -lms 100`)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self._p = None

    def __enter__(self):
        try:
            self._p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                        str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                       text=True)
            time.sleep(0.35)  # first sample before the timed region starts
        except Exception:
            self._p = None
        return self

    def __exit__(self, *a):
        if self._p is None:
            return
        time.sleep(0.15)
        self._p.terminate()
        try:
            out, _ = self._p.communicate(timeout=5)
        except Exception:
            self._p.kill()
            out = ""
        for line in (out or "").splitlines():
            cells = [c.strip() for c in line.split(",")]
            if len(cells) >= 9:
                self.rows.append(cells)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- reference arm
class CpuReference:
    """The reference's CPU implementation of the path on the host cores: oracle/_ref (the reference's own sources,
    compiled by oracle/Makefile where /root/reference exists) when present, else the oracle port (torch CPU)."""

    def __init__(self, model_name: str):
        self.model_name = model_name
        self.ref_bin = ROOT / "oracle" / "_ref" / "ref_decode_bench"
        self.kind = "reference" if self.ref_bin.exists() else "port"
        self.cores = 1
        self._state = None

    def _setup_port(self):
        import torch
        from oracle import decode_oracle as orc
        from tinygpt_b200 import models
        sys.path.insert(0, str(ROOT / "tests"))
        from helpers import to_oracle_cfg
        spec = models.SPECS[self.model_name].with_ctx(PROMPT_LEN + NEW_TOKENS + 16)
        w = {k: v.float() for k, v in models.synth_weights(spec, seed=0).items()}  # fp32 once, outside the clock
        cfg, table = to_oracle_cfg(spec), models.rope_table(spec)
        prompt = torch.randint(0, spec.vocab, (PROMPT_LEN,), generator=torch.Generator().manual_seed(0))
        cache = orc.KVCache()
        logits = orc.forward(cfg, w, prompt.view(1, -1), cache, table, "bf16")[:, -1]
        self.cores = torch.get_num_threads()
        self._state = (orc, cfg, w, table, cache, logits, spec)

    def step(self, n_tokens: int):
        """Decode n_tokens greedy tokens; returns (tokens/s, sample description)."""
        if self.kind == "reference":
            out = subprocess.run([str(self.ref_bin), "--model", self.model_name, "--prompt", str(PROMPT_LEN),
                                  "--tokens", str(n_tokens)], capture_output=True, text=True, timeout=1800)
            for line in out.stdout.splitlines():
                if line.startswith("{"):
                    d = json.loads(line)
                    self.cores = int(d.get("threads", 1))
                    return float(d["tokens_per_s"]), d.get("sample", "")
            print(f"[bench] oracle/_ref produced no result (rc={out.returncode}): {out.stderr[-400:]}; "
                  "falling back to the oracle port", file=sys.stderr)
            self.kind = "port"
        if self._state is None:
            self._setup_port()
        orc, cfg, w, table, cache, logits, spec = self._state
        if cache.past_length(0) + n_tokens > spec.max_ctx:
            self._setup_port()
            orc, cfg, w, table, cache, logits, spec = self._state
        t0 = time.perf_counter()
        for _ in range(n_tokens):
            tok = orc.argmax_last(logits)
            logits = orc.forward(cfg, w, tok.view(1, 1), cache, table, "bf16")[:, -1]
        dt = time.perf_counter() - t0
        self._state = (orc, cfg, w, table, cache, logits, spec)
        return n_tokens / dt, (f"{n_tokens} greedy decode steps continuing a {PROMPT_LEN}-token prompt "
                               "(oracle port: torch CPU fp32 matmul with the reference's bf16 rounding points)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = CpuReference(args.model)
    n_tokens = 4  # bounded sample per step: the CPU path runs at a few tokens/s at best
    vals, sample = [], ""
    t_start = time.perf_counter()
    for i in range(args.warmup + args.steps):
        v, sample = ref.step(n_tokens)
        if i >= args.warmup:
            vals.append(v)
        if time.perf_counter() - t_start > 240 and len(vals) >= 1:
            break  # keep the arm within a few minutes whatever K was asked for
    value = len(vals) / sum(1.0 / v for v in vals)  # total tokens / total time
    line = {
        "impl": "reference", "metric": "decode tokens/sec (bf16, batch=1)", "value": value, "unit": "tokens/s",
        "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup,
        "ms_per_step": 1000.0 * n_tokens / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{args.model} bf16 batch=1 {NEW_TOKENS}-token greedy decode after a {PROMPT_LEN}-token prompt",
                   "step": f"{n_tokens} decoded tokens (bounded sample of the {NEW_TOKENS}-token decode)"},
        "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": ref.cores, "kind": ref.kind, "sample": sample},
        "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------- our arm
def time_kernel_isolated(fn, iters=20, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def reference_cuda_baseline(spec, weights):
    """Decode tokens/s of the reference's own CUDA build (TinyTorch ops + cuBLAS + TinyFA, ≈ 490 launches per token) on
    this GPU, same model shape / synthetic weights / prompt length / 128 timed greedy steps.  Never fatal: any failure
    is reported as {"unavailable": why}.  Bounded: checkpoint write + load + 16-token prefill + 128 steps."""
    import tempfile
    bin_path = ROOT / "oracle" / "_ref" / "ref_cuda_decode"
    if not bin_path.exists():
        return None
    try:
        import torch
        sys.path.insert(0, str(ROOT / "tools"))
        import ref_cuda_parity as rp
        from tinygpt_b200 import models
        prompt = torch.randint(0, spec.vocab, (PROMPT_LEN,), generator=torch.Generator().manual_seed(0)).tolist()
        with tempfile.TemporaryDirectory() as td:
            models.save_checkpoint(spec, weights, td)
            # bounded: a first-ever run of this binary on a box must not be able to stretch the bench by minutes
            _, _, timing = rp.run_reference(spec, td, prompt, 2, time_steps=NEW_TOKENS, timeout=120)
            try:  # the same reference program with our engine behind GPTModel::model() (the TinyTorch adapter)
                _, _, t_adapter = rp.run_reference(spec, td, prompt, 2, time_steps=NEW_TOKENS, timeout=90, b200="engine")
            except Exception as e:  # noqa: BLE001
                t_adapter = {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}
        if not timing:
            return {"unavailable": "ref_cuda_decode printed no timing line"}
        out = {"value": timing["tokens_per_s"], "unit": "tokens/s", "us_per_token": timing["us_per_token"],
               "kind": "reference CUDA build (unmodified sources, oracle/_ref/ref_cuda_decode), same GPU",
               "sample": f"{NEW_TOKENS} greedy decode steps after a {PROMPT_LEN}-token prompt, device-resident token loop"}
        if t_adapter:
            out["same_program_with_b200_engine"] = (
                {"value": t_adapter["tokens_per_s"], "unit": "tokens/s", "us_per_token": t_adapter["us_per_token"],
                 "what": "the reference's own generate loop, our engine behind GPTModel::model() via "
                         "integration/tinytorch_b200_adapter.h"} if "tokens_per_s" in t_adapter else t_adapter)
        return out
    except Exception as e:  # noqa: BLE001 — a baseline must never take the bench line down
        return {"unavailable": f"{type(e).__name__}: {str(e)[:300]}"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from tinygpt_b200 import build, engine, models, ops
    from tinygpt_b200._lib import lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    build.build()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    spec = models.SPECS[args.model].with_ctx(PROMPT_LEN + NEW_TOKENS + 16)
    tp_mode = world > 1 and args.mode == "tp"
    # seeded synthetic checkpoint drawn on the GPU (same values on every rank: same seed, same Philox stream)
    w_full = models.synth_weights(spec, seed=0, device=dev, device_generator=True)
    if tp_mode:
        from tinygpt_b200 import tp
        eng = tp.TPDecodeEngine(spec, w_full, rank, world, dev)
        w = eng._w
    else:
        w = w_full
        eng = engine.DecodeEngine(spec, w)
    del w_full
    torch.cuda.empty_cache()
    # tensor parallel: every rank decodes the SAME sequence; replicas: one independent sequence per GPU
    prompt = torch.randint(0, spec.vocab, (1, PROMPT_LEN),
                           generator=torch.Generator().manual_seed(0 if tp_mode else rank)).to(dev)
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # prefill once; every timed step rewinds to the last prompt token and decodes NEW_TOKENS tokens from there
    # (NEW_TOKENS graph launches: the last prompt position is recomputed and yields token 0, then NEW_TOKENS-1 steps)
    eng.reset_cache()
    eng.gen_next_token(prompt)
    last_prompt_tok = prompt[:, -1:].contiguous()

    def step():
        eng.seek(PROMPT_LEN - 1)
        first = eng.gen_next_token(last_prompt_tok)
        return torch.cat([first.view(-1), eng.decode(NEW_TOKENS - 1)])

    for _ in range(max(args.warmup, 3)):
        toks = step()
    barrier()
    launches0 = lib().b200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            toks = step()
        e1.record(stream)
        barrier()
    ms = e0.elapsed_time(e1)
    launches = lib().b200_launch_count() - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    seqs = 1 if tp_mode else world                           # TP: one sequence on N GPUs; replicas: N sequences
    tokens_total = seqs * args.steps * NEW_TOKENS
    value = tokens_total / (ms_max / 1e3)
    ms_per_token = ms_max / (args.steps * NEW_TOKENS)

    # ---- e2e through the public API with host buffers
    pin_in = torch.empty(1, PROMPT_LEN, dtype=torch.int64).pin_memory()
    pin_out = torch.empty(NEW_TOKENS, dtype=torch.int64).pin_memory()
    prompt_host = prompt.cpu().view(-1).tolist()
    for _ in range(2):
        eng.generate_sync(prompt_host, NEW_TOKENS, pin_in, pin_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host_toks = eng.generate_sync(prompt_host, NEW_TOKENS, pin_in, pin_out)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = seqs * args.steps * NEW_TOKENS / float(t.item())
    # the host-buffer path (prompt through the batched GEMM prefill) and the resident path (last prompt token recomputed
    # by the decode kernels) agree up to bf16 near-ties of the synthetic model: same first token, ids in range
    ids_match = float((host_toks == toks.cpu()).float().mean())
    assert int(host_toks.min()) >= 0 and int(host_toks.max()) < spec.vocab

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline (rank 0)
    peak, peak_src = measured_peaks()
    ctx_mid = PROMPT_LEN + NEW_TOKENS // 2
    bytes_tok = eng.bytes_per_token(ctx_mid)                 # THIS rank's algorithmic bytes (1/N of the weights under TP)
    achieved = bytes_tok / (ms_per_token / 1e3) / 1e9
    # isolated timing of the two largest GEMV launches, live, CUDA events on the current stream; weights cycled so
    # that consecutive launches never hit L2 (lm_head 272 MB > L2; gate_up walks all layers: L × 17 MB)
    xh = torch.randn(spec.hidden, device=dev).to(torch.bfloat16)
    nw = w["model.norm.weight"]
    head = w["lm_head.weight"] if "lm_head.weight" in w else w["model.embed_tokens.weight"]
    t_head = time_kernel_isolated(lambda: ops.gemv_fused(xh, head, norm_weight=nw, eps=spec.rms_eps))
    li = [0]

    def gu():
        l = li[0] % spec.layers
        li[0] += 1
        ops.gemv_fused(xh, w[f"model.layers.{l}.mlp.gate_up_proj.weight"],
                       norm_weight=w[f"model.layers.{l}.post_attention_layernorm.weight"], eps=spec.rms_eps,
                       silu_mul=True)
    t_gu = time_kernel_isolated(gu, iters=2 * spec.layers, warm=spec.layers)
    head_bytes = 2 * head.shape[0] * spec.hidden
    gu_bytes = 2 * w["model.layers.0.mlp.gate_up_proj.weight"].shape[0] * spec.hidden
    kernels = [
        {"kernel": "gemv_stream_kernel lm_head (RMSNorm prologue)", "bytes": head_bytes, "ms": t_head,
         "achieved_gbs": head_bytes / t_head / 1e6, "frac": head_bytes / t_head / 1e6 / peak},
        {"kernel": "gemv_stream_kernel gate|up (RMSNorm prologue, SiLU·mul epilogue)", "bytes": gu_bytes, "ms": t_gu,
         "achieved_gbs": gu_bytes / t_gu / 1e6, "frac": gu_bytes / t_gu / 1e6 / peak},
    ]
    traffic = None
    tpath = ROOT / "profiles" / "traffic.json"
    if tpath.exists() and world == 1:
        try:
            traffic = json.loads(tpath.read_text()).get(spec.name)
        except Exception:
            traffic = None

    # ---- CPU baseline (bounded sample, rank 0, N = 1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        ref = CpuReference(args.model)
        v, sample = ref.step(8)
        cpu = {"value": v, "unit": "tokens/s", "cores": ref.cores, "kind": ref.kind, "sample": sample}

    # ---- the reference's OWN CUDA path on this GPU (a reported baseline like cpu_baseline; rank 0, N = 1 only): the
    # unmodified reference compiled from /root/reference into oracle/_ref/ref_cuda_decode (make -C oracle cuda), run
    # as a separate process on the same synthetic checkpoint shape, AFTER every measurement of ours
    ref_cuda = None
    if world == 1 and not args.no_cpu_baseline and os.environ.get("B200_BENCH_NO_REF_CUDA") != "1":
        ref_cuda = reference_cuda_baseline(spec, w)

    line = {
        "metric": "decode tokens/sec (bf16, batch=1)", "value": value, "unit": "tokens/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps,
        "ms_per_token": ms_per_token, "higher_is_better": True, "scaling": "strong" if tp_mode else "weak",
        "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{spec.name} bf16 batch=1 {NEW_TOKENS}-token greedy decode after a {PROMPT_LEN}-token prompt",
                   "step": f"seek({PROMPT_LEN - 1}) + {NEW_TOKENS} engine steps (one CUDA graph launch per token)",
                   "parallelism": "single GPU" if world == 1 else (
                       f"tp{world}: FFN columns + vocabulary sharded, attention heads " +
                       ("sharded" if eng.shard_attn else "replicated (head counts do not divide)") +
                       "; hidden-vector reductions fused into the GEMV kernels over NVLink peer memory"
                       if tp_mode else f"{world} independent replicas (one sequence per GPU, no collective)"),
                   "l2": (f"weights {eng.bytes_per_token(0) / 1e9:.2f} GB per GPU > 126 MB L2: re-streamed from HBM "
                          "every token, no explicit flush") if eng.bytes_per_token(0) > 126e6 else
                         (f"weights {eng.bytes_per_token(0) / 1e6:.0f} MB per GPU fit in the 126 MB L2 and are NOT "
                          "flushed between tokens: the HBM roofline does not bound this configuration"),
                   "kernel_sync": ("per-op completion counters (flag-sync)" if eng.options["flag_sync"] else
                                   "programmatic dependent launch (griddepcontrol.wait)"),
                   "l2_prefetch_mb": eng.options["l2_prefetch_mb"],
                   "gemv_smallk": os.environ.get("B200_GEMV_SMALLK", "default")},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "kernel": f"whole decode token ({eng.launches_per_token} launches, PDL-chained; "
                               "gemv_stream_kernel moves >99 % of the bytes)",
                     "bytes_per_launch": bytes_tok, "ms_per_launch": ms_per_token,
                     "frac_of_8TBs": achieved / 8000.0, "kernels": kernels},
        "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": PROMPT_LEN * 8,
                "d2h_bytes_per_step": NEW_TOKENS * 8,
                "what": f"generate_sync(host prompt) incl. H2D, reset, {PROMPT_LEN}-token prefill, decode, D2H",
                "ids_equal_to_resident_run": ids_match},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if ref_cuda is not None:
        line["reference_cuda_baseline"] = ref_cuda
    print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="Qwen2.5-0.5B")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="tp", choices=["tp", "replicas"],
                    help="N > 1: tensor parallel over one sequence (default, strong scaling) or independent replicas")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
