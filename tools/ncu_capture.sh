#!/bin/bash
# ncu --set full captures of the two largest GEMV launches of a Qwen2.5-0.5B decode token (run under gpurun).
set -u
mkdir -p gpurun_out
COMMON="--set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -f"
timeout 500 ncu $COMMON -k 'regex:gemv_stream_kernel<\(int\)4, \(int\)1, \(int\)1, \(int\)0>' -c 1 \
  -o gpurun_out/r01_lm_head python tools/profile_decode.py Qwen2.5-0.5B 1 > gpurun_out/ncu_full1.log 2>&1
tail -1 gpurun_out/ncu_full1.log
timeout 500 ncu $COMMON -k 'regex:gemv_stream_kernel<\(int\)1, \(int\)2, \(int\)1, \(int\)2>' -s 3 -c 2 \
  -o gpurun_out/r01_gate_up python tools/profile_decode.py Qwen2.5-0.5B 1 > gpurun_out/ncu_full2.log 2>&1
tail -1 gpurun_out/ncu_full2.log
timeout 500 ncu $COMMON -k 'regex:attn_decode_kernel' -s 3 -c 1 \
  -o gpurun_out/r01_attn python tools/profile_decode.py Qwen2.5-0.5B 1 > gpurun_out/ncu_full3.log 2>&1
tail -1 gpurun_out/ncu_full3.log
ls -la gpurun_out/*.ncu-rep
