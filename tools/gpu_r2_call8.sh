#!/bin/bash
# round-2 call 8: pair conv kernel with multi-step MMA asm blocks (the MMA warp was issue-bound: r02_c7 ncu), VAE rings kept
# across videos; parity, timing, ncu of the pair kernel at C = 96
mkdir -p gpurun_out
timeout 300 python tools/gpu_check.py conv > gpurun_out/r02_c8_check.log 2>&1
echo "check exit=$?"; grep -c "OK " gpurun_out/r02_c8_check.log; grep -E "BAD|rror" gpurun_out/r02_c8_check.log | head
timeout 300 python tools/gpu_check.py perf_conv > gpurun_out/r02_c8_perf_conv.log 2>&1
echo "perf_conv exit=$?"; grep PERF gpurun_out/r02_c8_perf_conv.log
timeout 300 python tools/vae_bench.py --iters 2 > gpurun_out/r02_c8_vae_bench.json 2> gpurun_out/r02_c8_vae_bench.err
echo "vae bench exit=$?"; tail -c 700 gpurun_out/r02_c8_vae_bench.json; echo
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv2_kernel -s 2 -c 1 -o gpurun_out/r02_c8_prof_conv2_96_mid -f \
    python tools/gpu_check.py perf_conv > gpurun_out/r02_c8_ncu_conv2_96.log 2>&1
echo "ncu conv2 96 exit=$?"
timeout 400 python -m pytest tests/test_vae_gpu.py tests/test_pipeline_gpu.py -m gpu -q -s > gpurun_out/r02_c8_pytest.log 2>&1
echo "pytest exit=$?"; grep -E "passed|failed|Error|BAD|levels" gpurun_out/r02_c8_pytest.log | tail -n 12
