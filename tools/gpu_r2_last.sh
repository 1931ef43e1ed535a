#!/bin/bash
# the round's last GPU minutes: q|k norm + RoPE kernel with its own loads issued before the partial-sum reduction (checks + timing)
mkdir -p gpurun_out
timeout 100 python tools/gpu_check.py abi3 > gpurun_out/r02_last_check.log 2>&1
echo "check exit=$?"; grep -c "OK " gpurun_out/r02_last_check.log; grep -E "BAD|rror" gpurun_out/r02_last_check.log | head
timeout 100 python tools/gpu_check.py perf_ew > gpurun_out/r02_last_perf_ew.log 2>&1
echo "perf exit=$?"; grep PERF gpurun_out/r02_last_perf_ew.log
