#!/bin/bash
# round-2 call 10: GEMM activation epilogue unswitched by hand again (ffn.0 regression of call 9), conv pair kernel with uniform
# MMA operands + accumulator released after the first epilogue pass; parity, timings, bench
mkdir -p gpurun_out
timeout 400 python tools/gpu_check.py conv abi3 gemm gemm_epi ln_fold > gpurun_out/r02_c10_check.log 2>&1
echo "check exit=$?"; grep -c "OK " gpurun_out/r02_c10_check.log; grep -E "BAD|rror" gpurun_out/r02_c10_check.log | head
timeout 300 python tools/gpu_check.py perf_conv > gpurun_out/r02_c10_perf_conv.log 2>&1
echo "perf_conv exit=$?"; grep PERF gpurun_out/r02_c10_perf_conv.log
timeout 300 python tools/gpu_check.py perf_gemm_epi > gpurun_out/r02_c10_perf_gemm.log 2>&1
echo "perf_gemm_epi exit=$?"; grep PERF gpurun_out/r02_c10_perf_gemm.log | head -20
timeout 300 python tools/vae_bench.py --iters 2 > gpurun_out/r02_c10_vae_bench.json 2> gpurun_out/r02_c10_vae_bench.err
echo "vae bench exit=$?"; tail -c 700 gpurun_out/r02_c10_vae_bench.json; echo
timeout 400 python bench.py --no-vae --breakdown > gpurun_out/r02_c10_bench.json 2> gpurun_out/r02_c10_bench.err
echo "bench exit=$?"; head -c 1700 gpurun_out/r02_c10_bench.json; echo; grep -E "e2e phases" gpurun_out/r02_c10_bench.err; grep -A 12 "breakdown of one step" gpurun_out/r02_c10_bench.err | head -n 14
