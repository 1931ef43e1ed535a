#!/bin/bash
# ncu evidence for profiles/ (run under gpurun on ONE GPU; numbers printed by these runs are never bench values):
#   1. launch list of one bench step  (--metrics gpu__time_duration.sum)  -> tools/launch_shares.py
#   2. full captures (--set full --import-source on) of: self-attention at L=32760, the pair GEMM with the o-projection
#      epilogue and with the GELU epilogue, the VAE conv                  -> tools/ncu_summary.py, tools/ncu_hot.py
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-vae > gpurun_out/ncu_bench_stdout.log 2>&1
echo "ncu launches exit=$?"; wc -l gpurun_out/launches.csv
# perf_attn: 7 launches at L=8192 without workspace + 7 with (each an attn_fwd launch), then L=32760
ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 16 -c 1 -o gpurun_out/prof_attn_final -f \
    python tools/gpu_check.py perf_attn > gpurun_out/ncu_attn_stdout.log 2>&1
echo "ncu attn exit=$?"
# perf_gemm_epi: 13 launches per case (3 warm + 10 timed): case 2 = o-projection, case 3 = ffn.0 + GELU
ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16 -s 16 -c 1 -o gpurun_out/prof_gemm_oproj -f \
    python tools/gpu_check.py perf_gemm_epi > gpurun_out/ncu_gemm_oproj_stdout.log 2>&1
echo "ncu gemm oproj exit=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16 -s 29 -c 1 -o gpurun_out/prof_gemm_gelu -f \
    python tools/gpu_check.py perf_gemm_epi > gpurun_out/ncu_gemm_gelu_stdout.log 2>&1
echo "ncu gemm gelu exit=$?"
ncu --set full --clock-control none --import-source on -k regex:conv_kernel -s 200 -c 1 -o gpurun_out/prof_conv_final -f \
    python tools/vae_bench.py --frames 17 --iters 1 > gpurun_out/ncu_conv_stdout.log 2>&1
echo "ncu conv exit=$?"; ls -la gpurun_out/*.ncu-rep
