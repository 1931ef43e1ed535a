#!/bin/bash
# round-2 call 7: convolution epilogue with batched residual loads + L2 prefetch (both kernels), pair kernel on narrow outputs;
# parity, timing, ncu launch list of a VAE round trip, full captures of the pair kernel at the two shapes that carry the FLOPs
mkdir -p gpurun_out
timeout 300 python tools/gpu_check.py conv abi3 > gpurun_out/r02_c7_check.log 2>&1
echo "check exit=$?"; grep -c "OK " gpurun_out/r02_c7_check.log; grep -E "BAD|rror" gpurun_out/r02_c7_check.log | head
timeout 300 python tools/gpu_check.py perf_conv > gpurun_out/r02_c7_perf_conv.log 2>&1
echo "perf_conv exit=$?"; grep PERF gpurun_out/r02_c7_perf_conv.log
timeout 300 python tools/vae_bench.py --iters 2 > gpurun_out/r02_c7_vae_bench.json 2> gpurun_out/r02_c7_vae_bench.err
echo "vae bench exit=$?"; tail -c 700 gpurun_out/r02_c7_vae_bench.json; echo
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_c7_vae_launches.csv \
    python tools/vae_bench.py --frames 17 --iters 1 > gpurun_out/r02_c7_ncu_vae_stdout.log 2>&1
echo "ncu vae launches exit=$?"; wc -l gpurun_out/r02_c7_vae_launches.csv
# perf_conv launches of conv2_kernel: 7 per (shape, mode); shape 0 = 96->96 (mid 0-6, end 7-13), shape 1 = 192->192 (mid 14-20, end 21-27)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv2_kernel -s 23 -c 1 -o gpurun_out/r02_c7_prof_conv2_192_end -f \
    python tools/gpu_check.py perf_conv > gpurun_out/r02_c7_ncu_conv2_192.log 2>&1
echo "ncu conv2 192 exit=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv2_kernel -s 2 -c 1 -o gpurun_out/r02_c7_prof_conv2_96_mid -f \
    python tools/gpu_check.py perf_conv > gpurun_out/r02_c7_ncu_conv2_96.log 2>&1
echo "ncu conv2 96 exit=$?"; ls -la gpurun_out/*.ncu-rep
timeout 600 python -m pytest tests/test_vae_gpu.py tests/test_harness_gpu.py -m gpu -q -s > gpurun_out/r02_c7_pytest.log 2>&1
echo "pytest exit=$?"; grep -E "passed|failed|Error|BAD|levels" gpurun_out/r02_c7_pytest.log | tail -n 12
