#!/bin/bash
# Runs tools/gpu_check.py sections in separate processes (a trapped kernel must not poison the rest).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for sec in "$@"; do
  timeout 300 python tools/gpu_check.py $sec > gpurun_out/check_$sec.log 2>&1
  echo "--- $sec exit=$? ---"
  tail -n 40 gpurun_out/check_$sec.log
done
