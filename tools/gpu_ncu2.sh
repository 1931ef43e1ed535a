#!/bin/bash
# full ncu captures: GEMM with the o-projection epilogue, GEMM with the GELU epilogue, attention at L=32760, VAE conv
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16 -s 16 -c 1 -o gpurun_out/prof_gemm_oproj -f \
    python tools/gpu_check.py perf_gemm_epi > gpurun_out/ncu_gemm_oproj_stdout.log 2>&1
echo "ncu gemm oproj exit=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16 -s 29 -c 1 -o gpurun_out/prof_gemm_gelu -f \
    python tools/gpu_check.py perf_gemm_epi > gpurun_out/ncu_gemm_gelu_stdout.log 2>&1
echo "ncu gemm gelu exit=$?"
ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 16 -c 1 -o gpurun_out/prof_attn_final -f \
    python tools/gpu_check.py perf_attn > gpurun_out/ncu_attn_stdout.log 2>&1
echo "ncu attn exit=$?"
ncu --set full --clock-control none --import-source on -k regex:conv_kernel -s 200 -c 1 -o gpurun_out/prof_conv_final -f \
    python tools/vae_bench.py --frames 17 --iters 1 > gpurun_out/ncu_conv_stdout.log 2>&1
echo "ncu conv exit=$?"; ls -la gpurun_out/*.ncu-rep
