#!/bin/bash
# round-2 call 9: pair conv kernel with one MMA asm block per operand stage + setmaxnreg role split with the residual held in
# registers; reproducible row sums of squares (stored partials instead of atomicAdd); parity, timing, determinism, bench
mkdir -p gpurun_out
timeout 400 python tools/gpu_check.py conv abi3 gemm_epi ln_fold ew > gpurun_out/r02_c9_check.log 2>&1
echo "check exit=$?"; grep -c "OK " gpurun_out/r02_c9_check.log; grep -E "BAD|rror" gpurun_out/r02_c9_check.log | head
timeout 300 python tools/gpu_check.py perf_conv > gpurun_out/r02_c9_perf_conv.log 2>&1
echo "perf_conv exit=$?"; grep PERF gpurun_out/r02_c9_perf_conv.log
timeout 300 python tools/vae_bench.py --iters 2 > gpurun_out/r02_c9_vae_bench.json 2> gpurun_out/r02_c9_vae_bench.err
echo "vae bench exit=$?"; tail -c 700 gpurun_out/r02_c9_vae_bench.json; echo
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv2_kernel -s 9 -c 1 -o gpurun_out/r02_c9_prof_conv2_96_end -f \
    python tools/gpu_check.py perf_conv > gpurun_out/r02_c9_ncu_conv2_96.log 2>&1
echo "ncu conv2 96 end exit=$?"
timeout 900 python -m pytest tests/test_dit_gpu.py tests/test_vae_gpu.py tests/test_pipeline_gpu.py tests/test_harness_gpu.py -m gpu -q -s > gpurun_out/r02_c9_pytest.log 2>&1
echo "pytest exit=$?"; grep -E "passed|failed|Error|BAD|levels|inside" gpurun_out/r02_c9_pytest.log | tail -n 24
timeout 400 python bench.py --no-vae --breakdown > gpurun_out/r02_c9_bench.json 2> gpurun_out/r02_c9_bench.err
echo "bench exit=$?"; head -c 2000 gpurun_out/r02_c9_bench.json; echo; grep -E "e2e phases" gpurun_out/r02_c9_bench.err; grep -A 18 "breakdown of one step" gpurun_out/r02_c9_bench.err | head -n 22
