#!/bin/bash
# ncu evidence: (1) launch list with per-launch device time for one bench step, (2) full capture of the top kernel
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench_stdout.log 2>&1
echo "ncu launches exit=$?"; wc -l gpurun_out/launches.csv
ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 4 -c 1 -o gpurun_out/prof_attn -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_attn_stdout.log 2>&1
echo "ncu attn exit=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 40 -c 1 -o gpurun_out/prof_gemm -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_gemm_stdout.log 2>&1
echo "ncu gemm exit=$?"; ls -la gpurun_out/*.ncu-rep
