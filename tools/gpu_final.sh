#!/bin/bash
# one call: kernel checks, row-kernel perf, the whole GPU test suite, the bench line (+ breakdown), the launch list
mkdir -p gpurun_out
bash tools/gpu_sections.sh ew perf_ew 2>&1 | grep -v "^\[OK"
grep -c "OK" gpurun_out/check_ew.log
bash tools/gpu_tests.sh
bash tools/gpu_bench.sh --steps 3 --warmup 3 --breakdown
sed -n '/breakdown of one step/,$p' gpurun_out/bench.err | head -12
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-vae > gpurun_out/ncu_bench_stdout.log 2>&1
echo "ncu launches exit=$?"; wc -l gpurun_out/launches.csv
