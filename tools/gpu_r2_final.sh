#!/bin/bash
# round-2 final single-GPU verification: the whole GPU suite (the driver's command), smoke(), the default bench line, and the ncu
# launch list of one bench step (kernel shares for profiles/)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r02_final_pytest.log 2>&1
echo "pytest exit=$?"; grep -E "passed|failed|rror|BAD" gpurun_out/r02_final_pytest.log | tail -n 12
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" > gpurun_out/r02_final_smoke.log 2>&1
echo "smoke exit=$?"; tail -n 3 gpurun_out/r02_final_smoke.log
timeout 900 python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err
echo "bench exit=$?"; head -c 3500 gpurun_out/r02_final_bench.json; echo; grep -E "e2e phases" gpurun_out/r02_final_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_final_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-vae > gpurun_out/r02_final_ncu_bench_stdout.log 2>&1
echo "ncu launches exit=$?"; wc -l gpurun_out/r02_final_launches.csv
