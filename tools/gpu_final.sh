#!/bin/bash
# one call: the whole GPU test suite, the bench line, flagship-model and 720p VAE points, final ncu captures
mkdir -p gpurun_out
bash tools/gpu_tests.sh
bash tools/gpu_bench.sh --steps 3 --warmup 3
timeout 600 python tools/dit14b_bench.py > gpurun_out/dit14b_bench.json 2> gpurun_out/dit14b_bench.err; echo "dit14b exit=$?"; tail -2 gpurun_out/dit14b_bench.err; cat gpurun_out/dit14b_bench.json
timeout 600 python tools/vae_bench.py --frames 81 --height 720 --width 1280 --iters 1 > gpurun_out/vae_bench_720p.json 2> gpurun_out/vae_bench_720p.err; echo "vae720 exit=$?"; tail -2 gpurun_out/vae_bench_720p.err; cat gpurun_out/vae_bench_720p.json
ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 16 -c 1 -o gpurun_out/prof_attn_final -f \
    python tools/gpu_check.py perf_attn > gpurun_out/ncu_attn_stdout.log 2>&1
echo "ncu attn exit=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16 -s 16 -c 1 -o gpurun_out/prof_gemm_oproj -f \
    python tools/gpu_check.py perf_gemm_epi > gpurun_out/ncu_gemm_oproj_stdout.log 2>&1
echo "ncu gemm oproj exit=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16 -s 29 -c 1 -o gpurun_out/prof_gemm_gelu -f \
    python tools/gpu_check.py perf_gemm_epi > gpurun_out/ncu_gemm_gelu_stdout.log 2>&1
echo "ncu gemm gelu exit=$?"
