#!/bin/bash
# DRAM bytes of every kernel of one decode token (cheap metric pass) → gpurun_out/token_traffic_<model>.csv
set -u
M=${1:-Qwen2.5-0.5B}
mkdir -p gpurun_out
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  --profile-from-start off --csv --log-file gpurun_out/token_traffic_$M.csv python tools/profile_decode.py $M 1 > /dev/null 2>&1
wc -l gpurun_out/token_traffic_$M.csv
