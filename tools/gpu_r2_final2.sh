#!/bin/bash
# round-2 last single-GPU call: q / k norm kernel with the warp-cooperative reduction of the partial sums (kernel checks + DiT
# parity), bench (e2e per-step wall times), ncu launch list of one bench step with the model build skipped
mkdir -p gpurun_out
timeout 300 python tools/gpu_check.py abi3 ew > gpurun_out/r02_f2_check.log 2>&1
echo "check exit=$?"; grep -c "OK " gpurun_out/r02_f2_check.log; grep -E "BAD|rror" gpurun_out/r02_f2_check.log | head
timeout 600 python -m pytest tests/test_dit_gpu.py -m gpu -q -s > gpurun_out/r02_f2_pytest.log 2>&1
echo "pytest exit=$?"; grep -E "passed|failed|rror|inside|twice|files vs" gpurun_out/r02_f2_pytest.log | tail -n 16
timeout 600 python bench.py --no-vae --breakdown > gpurun_out/r02_f2_bench.json 2> gpurun_out/r02_f2_bench.err
echo "bench exit=$?"; head -c 2600 gpurun_out/r02_f2_bench.json; echo; grep -E "e2e phases|e2e step wall" gpurun_out/r02_f2_bench.err; grep -A 10 "breakdown of one step" gpurun_out/r02_f2_bench.err | head -n 12
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02_f2_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-vae > gpurun_out/r02_f2_ncu_bench_stdout.log 2>&1
echo "ncu launches exit=$?"; wc -l gpurun_out/r02_f2_launches.csv
