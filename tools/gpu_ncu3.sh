#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16 -s 16 -c 1 -o gpurun_out/prof_gemm_oproj2 -f \
    python tools/gpu_check.py perf_gemm_epi > gpurun_out/ncu_gemm_oproj_stdout.log 2>&1
echo "ncu gemm oproj exit=$?"
