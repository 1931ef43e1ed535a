#!/bin/bash
# round-2 call 5: column-split softmax variant of the attention kernel (correctness, then timing), GEMM with the stage count
# chosen per epilogue class, bench with the e2e host profile
mkdir -p gpurun_out
LIBDIR=$PWD/stable-video-infinity_b200/lib
SVI_B200_LIB=$LIBDIR/libsvi_b200_attn_colsplit.so timeout 300 python tools/gpu_check.py attn attn_cross abi3 attn_bench > gpurun_out/r02_c5_check_colsplit.log 2>&1
echo "colsplit check exit=$?"; grep -c "OK " gpurun_out/r02_c5_check_colsplit.log; grep -E "BAD|timeout|Error" gpurun_out/r02_c5_check_colsplit.log | head -10
for v in "" _attn_colsplit _attn_colsplit_p4 _attn_colsplit_p7 _attn_colsplit_r192 _attn_r1; do
  SVI_B200_LIB=$LIBDIR/libsvi_b200$v.so timeout 120 python tools/gpu_check.py perf_attn_quick 2>&1 | grep PERF
done | tee gpurun_out/r02_c5_attn_variants.log
timeout 300 python tools/gpu_check.py gemm gemm_epi ln_fold > gpurun_out/r02_c5_check.log 2>&1
echo "check exit=$?"; grep -c "OK " gpurun_out/r02_c5_check.log; grep -E "BAD|Error|error" gpurun_out/r02_c5_check.log | head
timeout 300 python tools/gpu_check.py perf_gemm_epi > gpurun_out/r02_c5_perf.log 2>&1
echo "perf exit=$?"; grep PERF gpurun_out/r02_c5_perf.log
timeout 600 python -m pytest tests/test_dit_gpu.py tests/test_kernels_gpu.py -m gpu -q -s > gpurun_out/r02_c5_pytest.log 2>&1
echo "pytest exit=$?"; grep -E "passed|failed|Error|BAD" gpurun_out/r02_c5_pytest.log | tail -n 12
SVI_BENCH_PROFILE=1 timeout 400 python bench.py --no-vae > gpurun_out/r02_c5_bench.json 2> gpurun_out/r02_c5_bench.err
echo "bench exit=$?"; head -c 2300 gpurun_out/r02_c5_bench.json; echo; grep -E "e2e phases" gpurun_out/r02_c5_bench.err; grep -A 26 "cumulative" gpurun_out/r02_c5_bench.err | head -n 40
SVI_B200_LIB=$LIBDIR/libsvi_b200_attn_colsplit.so timeout 400 python bench.py --no-vae --no-cpu-baseline --no-e2e > gpurun_out/r02_c5_bench_colsplit.json 2> gpurun_out/r02_c5_bench_colsplit.err
echo "bench colsplit exit=$?"; head -c 1900 gpurun_out/r02_c5_bench_colsplit.json; echo
