#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:attn -s 7 -c 1 -o gpurun_out/prof_attn_${1:-v2} -f \
    python tools/gpu_check.py perf_attn > gpurun_out/ncu_attn_${1:-v2}_stdout.log 2>&1
echo "ncu exit=$?"; ls -la gpurun_out/prof_attn_${1:-v2}.ncu-rep
