mkdir -p gpurun_out/c22
timeout 400 python -m pytest tests/test_batch_gpu.py -q -m gpu -s 2>&1 | grep -v "^$" | tail -40 > gpurun_out/c22/tests.log
tail -15 gpurun_out/c22/tests.log
for k in mma exact; do echo "== B200_BATCH_GEMV=$k"; B200_BATCH_GEMV=$k timeout 200 python tools/batch_bench.py Qwen2.5-0.5B Llama-3.2-3B Mistral-7B-v0.3 2>&1; done | tee gpurun_out/c22/batch_bench.txt
