#!/bin/bash
mkdir -p gpurun_out
for d in 2 3 4; do
  SVI_ATTN_DEBUG=$d timeout 200 python tools/gpu_check.py perf_attn 2>&1 | grep -E "PERF\] attn L=32760|PERF\] cross" | sed "s/^/[DBG=$d] /"
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv -lms 100 > gpurun_out/clk_attn.csv &
SMI=$!
timeout 200 python tools/gpu_check.py perf_attn 2>&1 | grep -E "PERF\] attn"
kill $SMI
sort gpurun_out/clk_attn.csv | uniq -c | sort -rn | head -5
