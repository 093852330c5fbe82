#!/usr/bin/env bash
# One GPU call that validates and measures everything that was STAGED without hardware:
#   gpurun --timeout 2400 -- 'bash tools/staged_gpu.sh'            (everything: ~35 GPU-minutes)
#   gpurun --timeout 900  -- 'bash tools/staged_gpu.sh tests ref sync'   (correctness first: ~10 GPU-minutes)
# Results land in gpurun_out/staged/.  Nothing here changes a default; each block is independent (a failure does not
# stop the next one) and bounded by its own timeout so a hang cannot eat the call.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/staged
mkdir -p "$OUT"
python -m tinygpt_b200.build > "$OUT/build.log" 2>&1
# sections: tests ref sync decode smallk l2pf prefill gemm   (default: all)   e.g.  bash tools/staged_gpu.sh tests sync decode
WANT=" ${*:-all} "
want() { [[ "$WANT" == *" all "* || "$WANT" == *" $1 "* ]]; }

if want tests; then
echo "== staged tests"
# one process per group: a trap (sticky CUDA error) in one staged path must not mask the others
: > "$OUT/staged_tests.log"
for grp in flagsync_engine flagsync_full mma_causal mma_prefill persistent_gemm prefill_chunk generate_async loader_cuda \
           l2_prefetch smallk_gemv smallk_engine tp2_flagsync reference_cuda drop_in_boundary device_sampler engine_sampler; do
  echo "---- $grp" >> "$OUT/staged_tests.log"
  B200_STAGED=1 timeout 600 python -m pytest tests/test_staged_gpu.py -m gpu -q -rA -s -k "$grp" >> "$OUT/staged_tests.log" 2>&1
  echo "$grp: $(grep -E '^[0-9]+ (passed|failed)|passed|failed|error' "$OUT/staged_tests.log" | tail -n 1)"
done
fi

if want ref; then
echo "== engine and oracle against the reference's own CUDA path (oracle/_ref/ref_cuda_decode, if it was built)"
if [ -x oracle/_ref/ref_cuda_decode ]; then
  timeout 2400 python tools/ref_cuda_parity.py --prefill Qwen3-1.7B:2048 --json "$OUT/ref_cuda_parity.json" > "$OUT/ref_cuda_parity.log" 2>&1
  tail -n 12 "$OUT/ref_cuda_parity.log"
else
  echo "not built: run 'make -C oracle cuda' in the container that has /root/reference"
fi
fi

if want sync; then
echo "== grid-dependency microbenchmark"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/sync_bench tools/micro/sync_bench.cu > "$OUT/sync_bench.log" 2>&1 \
  && timeout 120 /tmp/sync_bench >> "$OUT/sync_bench.log" 2>&1
cat "$OUT/sync_bench.log"
fi

if want decode; then
echo "== decode: PDL (default) vs flag counters"
for m in Qwen2.5-0.5B Llama-3.2-3B; do
  timeout 400 python bench.py --model $m --steps 10 --warmup 3 --no-cpu-baseline > "$OUT/bench_${m}_pdl.json" 2> "$OUT/bench_${m}_pdl.err"
  B200_FLAGSYNC=1 timeout 400 python bench.py --model $m --steps 10 --warmup 3 --no-cpu-baseline > "$OUT/bench_${m}_flagsync.json" 2> "$OUT/bench_${m}_flagsync.err"
  python - "$OUT/bench_${m}_pdl.json" "$OUT/bench_${m}_flagsync.json" <<'PY'
import json, sys
for p in sys.argv[1:]:
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
        print(f"{p}: {d['value']:.0f} tok/s  {d['ms_per_token']*1e3:.1f} us/token  frac {d['roofline']['frac']:.3f}  e2e {d['e2e']['value']:.0f}")
    except Exception as e:
        print(p, "no result:", e)
PY
done
fi

if want smallk; then
echo "== decode 0.5B: register-resident small-k loop, alone and with flag counters"
B200_GEMV_SMALLK=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > "$OUT/bench_Qwen2.5-0.5B_smallk.json" 2> "$OUT/bench_Qwen2.5-0.5B_smallk.err"
B200_GEMV_SMALLK=1 B200_FLAGSYNC=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > "$OUT/bench_Qwen2.5-0.5B_smallk_flagsync.json" 2> "$OUT/bench_Qwen2.5-0.5B_smallk_flagsync.err"
python - "$OUT"/bench_Qwen2.5-0.5B_smallk*.json <<'PY'
import json, sys
for p in sys.argv[1:]:
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
        print(f"{p}: {d['value']:.0f} tok/s  {d['ms_per_token']*1e3:.1f} us/token  frac {d['roofline']['frac']:.3f}")
    except Exception as e:
        print(p, "no result:", e)
PY
B200_GEMV_SMALLK=1 timeout 200 python tools/trace_decode.py Qwen2.5-0.5B > "$OUT/trace_smallk.log" 2>&1
tail -n 8 "$OUT/trace_smallk.log"
fi

if want l2pf; then
echo "== decode: cross-kernel L2 prefetch (hint only) on the HBM-bound models"
for m in Llama-3.2-3B Mistral-7B-v0.3; do
  for mb in 8 24; do
    B200_L2PF_MB=$mb timeout 400 python bench.py --model $m --steps 10 --warmup 3 --no-cpu-baseline > "$OUT/bench_${m}_l2pf${mb}.json" 2> "$OUT/bench_${m}_l2pf${mb}.err"
  done
  B200_L2PF_MB=24 B200_FLAGSYNC=1 timeout 400 python bench.py --model $m --steps 10 --warmup 3 --no-cpu-baseline > "$OUT/bench_${m}_l2pf24_flagsync.json" 2> "$OUT/bench_${m}_l2pf24_flagsync.err"
done
timeout 400 python bench.py --model Mistral-7B-v0.3 --steps 10 --warmup 3 --no-cpu-baseline > "$OUT/bench_Mistral-7B-v0.3_pdl.json" 2> "$OUT/bench_Mistral-7B-v0.3_pdl.err"
python - "$OUT"/bench_*l2pf*.json "$OUT/bench_Mistral-7B-v0.3_pdl.json" <<'PY'
import json, sys
for p in sys.argv[1:]:
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
        print(f"{p}: {d['value']:.0f} tok/s  {d['ms_per_token']*1e3:.1f} us/token  frac {d['roofline']['frac']:.3f}")
    except Exception as e:
        print(p, "no result:", e)
PY
B200_FLAGSYNC=1 timeout 200 python tools/trace_decode.py Qwen2.5-0.5B > "$OUT/trace_flagsync.log" 2>&1
timeout 200 python tools/trace_decode.py Qwen2.5-0.5B > "$OUT/trace_pdl.log" 2>&1
tail -n 8 "$OUT/trace_pdl.log" "$OUT/trace_flagsync.log"
fi

if want prefill; then
echo "== prefill (config 4): default / mma attention / + persistent GEMM / + one 2048-token chunk"
timeout 300 python tools/prefill_bench.py > "$OUT/prefill_default.log" 2>&1
B200_PREFILL_ATTN=mma timeout 300 python tools/prefill_bench.py > "$OUT/prefill_mma.log" 2>&1
B200_PREFILL_ATTN=mma B200_GEMM=persistent timeout 300 python tools/prefill_bench.py > "$OUT/prefill_mma_pgemm.log" 2>&1
B200_PREFILL_ATTN=mma B200_GEMM=persistent B200_PREFILL_CHUNK=2048 timeout 300 python tools/prefill_bench.py > "$OUT/prefill_mma_pgemm_chunk2048.log" 2>&1
tail -n 1 "$OUT"/prefill_*.log
fi

if want gemm; then
echo "== GEMM alone"
timeout 300 python tools/gemm_bench.py > "$OUT/gemm_bench.log" 2>&1
cat "$OUT/gemm_bench.log"
fi
