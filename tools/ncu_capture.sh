#!/bin/bash
# ncu --set full captures of the kernels of a Qwen2.5-0.5B decode token that matter (run under gpurun):
#   ROUND=r02 bash tools/ncu_capture.sh  → gpurun_out/r02_{lm_head,gate_up,qkv,down,attn}.ncu-rep
# ncu serialises the launches and replays each ≈ 40 times: no PDL overlap, cold caches — durations are upper bounds, the
# stall reasons on the source page are what these captures are for.
set -u
ROUND=${ROUND:-r02}
MODEL=${MODEL:-Qwen2.5-0.5B}
mkdir -p gpurun_out
COMMON="--set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -f"
cap() {  # name, kernel regex, skip
  timeout 500 ncu $COMMON -k "regex:$2" -s "$3" -c 1 -o gpurun_out/${ROUND}_$1 python tools/profile_decode.py "$MODEL" 1 \
    > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log
}
cap lm_head 'gemv_stream_kernel<\(int\)4, \(int\)1, \(int\)1, \(int\)0' 0
cap gate_up 'gemv_(xreg|stream)_kernel<(\(int\)1, )?\(int\)2, \(int\)1, \(int\)2' 3
cap qkv 'gemv_(xreg|stream)_kernel<(\(int\)1, )?\(int\)1, \(int\)1, \(int\)0' 3
cap down 'gemv_stream_kernel<\(int\)1, \(int\)1, \(int\)0, \(int\)1' 3
cap attn 'attn_decode' 3
ls -la gpurun_out/*.ncu-rep
