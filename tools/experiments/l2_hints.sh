# L2 eviction hints (B200_HINTS: bit 0 = norm weights evict_last, bit 1 = weight stream evict_first): batched and batch-1 steps
mkdir -p gpurun_out/c25
for h in 0 1 2 3; do echo "== B200_HINTS=$h"; B200_HINTS=$h timeout 150 python tools/batch_bench.py Qwen2.5-0.5B Mistral-7B-v0.3 2>&1 | cut -c1-75; done | tee gpurun_out/c25/hints.txt
for h in 1 3; do echo "== trace B200_HINTS=$h"; B200_HINTS=$h timeout 90 python tools/trace_decode.py Qwen2.5-0.5B 8 2>&1 | tail -8; B200_HINTS=$h timeout 90 python tools/trace_decode.py Qwen2.5-0.5B 1 2>&1 | tail -8; done | tee gpurun_out/c25/hints_trace.txt
