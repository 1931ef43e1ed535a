#!/bin/bash
# Runs tools/gpu_check.py sections in separate processes (a trapped kernel must not poison the rest).
# LOGSUFFIX=<tag> keeps the logs of A/B runs apart.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for sec in "$@"; do
  timeout 300 python tools/gpu_check.py $sec > gpurun_out/check_${sec}${LOGSUFFIX}.log 2>&1
  echo "--- $sec${LOGSUFFIX} exit=$? ---"
  tail -n 40 gpurun_out/check_${sec}${LOGSUFFIX}.log
done
