#!/bin/bash
# round-2 call 1: full GPU suite + smoke + bench (with parity) + fast-math A/B + step breakdown + kernel perf
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_c1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -s -x > gpurun_out/r02_c1_pytest.log 2>&1
echo "pytest exit=$?"
grep -E "inside=|passed|failed|Error|error|BAD" gpurun_out/r02_c1_pytest.log | tail -n 60
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c1_smoke.log 2>&1
echo "smoke exit=$?"; tail -n 3 gpurun_out/r02_c1_smoke.log
timeout 600 python bench.py > gpurun_out/r02_c1_bench.json 2> gpurun_out/r02_c1_bench.err
echo "bench exit=$?"; cat gpurun_out/r02_c1_bench.json | head -c 3000; tail -n 5 gpurun_out/r02_c1_bench.err
SVI_B200_LIB=$PWD/stable-video-infinity_b200/lib/libsvi_b200_fastmath.so timeout 400 python bench.py --no-cpu-baseline --no-vae > gpurun_out/r02_c1_bench_fastmath.json 2> gpurun_out/r02_c1_bench_fastmath.err
echo "bench fastmath exit=$?"; head -c 1500 gpurun_out/r02_c1_bench_fastmath.json
timeout 400 python bench.py --breakdown --no-cpu-baseline --no-vae --no-e2e > gpurun_out/r02_c1_bench_bd.json 2> gpurun_out/r02_c1_breakdown.txt
echo "breakdown exit=$?"; cat gpurun_out/r02_c1_breakdown.txt | tail -n 30
timeout 300 python tools/gpu_check.py perf_attn perf_gemm_epi perf_ew > gpurun_out/r02_c1_perf.log 2>&1
echo "perf exit=$?"; grep PERF gpurun_out/r02_c1_perf.log
