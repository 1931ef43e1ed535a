#!/bin/bash
# ncu evidence for profiles/: launch list of one bench step + full captures of the three tensor kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench_stdout.log 2>&1
echo "ncu launches exit=$?"; wc -l gpurun_out/launches.csv
ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 9 -c 1 -o gpurun_out/prof_attn_final -f \
    python tools/gpu_check.py perf_attn > gpurun_out/ncu_attn_stdout.log 2>&1
echo "ncu attn exit=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 28 -c 1 -o gpurun_out/prof_gemm_final -f \
    python tools/gpu_check.py perf_gemm > gpurun_out/ncu_gemm_stdout.log 2>&1
echo "ncu gemm exit=$?"
ncu --set full --clock-control none --import-source on -k regex:conv_kernel -s 200 -c 1 -o gpurun_out/prof_conv_final -f \
    python tools/vae_bench.py --frames 17 --iters 1 > gpurun_out/ncu_conv_stdout.log 2>&1
echo "ncu conv exit=$?"; ls -la gpurun_out/*final.ncu-rep
