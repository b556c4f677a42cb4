#!/bin/bash
# Per-kernel DRAM traffic of ONE step (tools/step_for_ncu.py), fp32-parity and bf16 engines:
#   cold  = ncu's default cache control (L2 flushed before every kernel: reads are worst case, write-backs mostly fall
#           outside the kernel's own window)
#   warm  = --cache-control none (what the step really moves: producers leave their output in the 126 MB L2)
# Summarise with profiles/summarize_traffic.py.  Run on the GPU box:  gpurun -- 'bash tools/ncu_traffic.sh'
set -e
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
mkdir -p gpurun_out
for P in fp32 bf16; do
  timeout 300 ncu --profile-from-start off --metrics $M --clock-control none --csv \
      --log-file gpurun_out/traffic_cold_$P.csv python tools/step_for_ncu.py --precision $P > gpurun_out/traffic_cold_$P.log 2>&1
  timeout 300 ncu --profile-from-start off --metrics $M --clock-control none --cache-control none --csv \
      --log-file gpurun_out/traffic_warm_$P.csv python tools/step_for_ncu.py --precision $P > gpurun_out/traffic_warm_$P.log 2>&1
done
echo done
