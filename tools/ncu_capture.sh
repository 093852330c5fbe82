#!/bin/bash
# ncu --set full captures of the two largest GEMV launches and the attention of a Qwen2.5-0.5B decode token (run under
# gpurun).  ROUND=r02 bash tools/ncu_capture.sh  → gpurun_out/r02_{lm_head,gate_up,attn}.ncu-rep
# Kernel-name regexes stop before the trailing template flags (…, (bool)FS, (bool)SMALLK>), so they match the PDL, the
# flag-sync and the small-k instantiations alike: whatever the engine was built with is what gets captured.
set -u
ROUND=${ROUND:-r02}
mkdir -p gpurun_out
COMMON="--set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -f"
timeout 500 ncu $COMMON -k 'regex:gemv_stream_kernel<\(int\)4, \(int\)1, \(int\)1, \(int\)0' -c 1 \
  -o gpurun_out/${ROUND}_lm_head python tools/profile_decode.py Qwen2.5-0.5B 1 > gpurun_out/ncu_full1.log 2>&1
tail -1 gpurun_out/ncu_full1.log
timeout 500 ncu $COMMON -k 'regex:gemv_stream_kernel<\(int\)1, \(int\)2, \(int\)1, \(int\)2' -s 3 -c 2 \
  -o gpurun_out/${ROUND}_gate_up python tools/profile_decode.py Qwen2.5-0.5B 1 > gpurun_out/ncu_full2.log 2>&1
tail -1 gpurun_out/ncu_full2.log
timeout 500 ncu $COMMON -k 'regex:attn_decode_kernel' -s 3 -c 1 \
  -o gpurun_out/${ROUND}_attn python tools/profile_decode.py Qwen2.5-0.5B 1 > gpurun_out/ncu_full3.log 2>&1
tail -1 gpurun_out/ncu_full3.log
ls -la gpurun_out/*.ncu-rep
