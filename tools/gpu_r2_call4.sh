#!/bin/bash
# round-2 call 4: attention final sweep; coalesced residual epilogue + LN fold checks and microbench; bench fold on/off;
# BASELINE configs[4] (VAE sweep) and configs[3] (clip loop, 1 GPU, 3 clips); ncu captures of the o-projection GEMM and the conv
mkdir -p gpurun_out
LIBDIR=$PWD/stable-video-infinity_b200/lib
for v in "" _attn_poly4 _attn_poly6 _attn_regs200 _attn_regs224 _attn_r1; do
  SVI_B200_LIB=$LIBDIR/libsvi_b200$v.so timeout 120 python tools/gpu_check.py perf_attn_quick 2>&1 | grep PERF
done | tee gpurun_out/r02_c4_attn_variants.log
timeout 300 python tools/gpu_check.py gemm gemm_epi ln_fold attn > gpurun_out/r02_c4_check.log 2>&1
echo "check exit=$?"; grep -c "OK " gpurun_out/r02_c4_check.log; grep -E "BAD|Error|error" gpurun_out/r02_c4_check.log | head -20; grep "fold" gpurun_out/r02_c4_check.log | head -20
timeout 300 python tools/gpu_check.py perf_gemm_epi perf_ln_fold > gpurun_out/r02_c4_perf.log 2>&1
echo "perf exit=$?"; grep PERF gpurun_out/r02_c4_perf.log
timeout 600 python -m pytest tests/test_dit_gpu.py tests/test_kernels_gpu.py -m gpu -q -s > gpurun_out/r02_c4_pytest.log 2>&1
echo "pytest exit=$?"; grep -E "inside=|passed|failed|Error|BAD" gpurun_out/r02_c4_pytest.log | tail -n 24
timeout 400 python bench.py --no-vae > gpurun_out/r02_c4_bench.json 2> gpurun_out/r02_c4_bench.err
echo "bench exit=$?"; head -c 2300 gpurun_out/r02_c4_bench.json; echo; grep -E "e2e phases" gpurun_out/r02_c4_bench.err
SVI_LN_FOLD=0 timeout 400 python bench.py --no-vae --no-cpu-baseline --no-e2e --breakdown > gpurun_out/r02_c4_bench_nofold.json 2> gpurun_out/r02_c4_bench_nofold.err
echo "bench nofold exit=$?"; head -c 500 gpurun_out/r02_c4_bench_nofold.json; echo; tail -n 18 gpurun_out/r02_c4_bench_nofold.err
timeout 400 python bench.py --no-vae --no-cpu-baseline --no-e2e --breakdown > gpurun_out/r02_c4_bench_fold_bd.json 2> gpurun_out/r02_c4_bench_fold_bd.err
echo "bench fold breakdown exit=$?"; tail -n 18 gpurun_out/r02_c4_bench_fold_bd.err
timeout 500 python bench.py --workload cfg5 > gpurun_out/r02_c4_cfg5.json 2> gpurun_out/r02_c4_cfg5.err
echo "cfg5 exit=$?"; head -c 3000 gpurun_out/r02_c4_cfg5.json; echo; tail -n 3 gpurun_out/r02_c4_cfg5.err
timeout 500 python bench.py --workload cfg4 --clips 3 > gpurun_out/r02_c4_cfg4.json 2> gpurun_out/r02_c4_cfg4.err
echo "cfg4 exit=$?"; head -c 1500 gpurun_out/r02_c4_cfg4.json; echo; tail -n 3 gpurun_out/r02_c4_cfg4.err
ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16 -s 16 -c 1 -o gpurun_out/r02_prof_gemm_oproj -f \
    python tools/gpu_check.py perf_gemm_epi > gpurun_out/r02_c4_ncu_gemm_stdout.log 2>&1
echo "ncu gemm exit=$?"
ncu --set full --clock-control none --import-source on -k regex:conv_kernel -s 200 -c 1 -o gpurun_out/r02_prof_conv -f \
    python tools/vae_bench.py --frames 17 --iters 1 > gpurun_out/r02_c4_ncu_conv_stdout.log 2>&1
echo "ncu conv exit=$?"; ls -la gpurun_out/r02_prof_*.ncu-rep
