#!/bin/bash
# round-2 call 2: new attention kernel (setmaxnreg, 3/16 polynomial share) checked first; falls back to the round-1 kernel
# library for the rest of the call if it fails, so the other measurements are not lost.
mkdir -p gpurun_out
LIBDIR=$PWD/stable-video-infinity_b200/lib
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_c2_smi.txt 2>&1
timeout 300 python tools/gpu_check.py attn attn_cross abi3 > gpurun_out/r02_c2_check_attn.log 2>&1
rc=$?
echo "attn check exit=$rc"; grep -c "OK " gpurun_out/r02_c2_check_attn.log; grep -E "BAD|timeout|Error" gpurun_out/r02_c2_check_attn.log | head -10
if [ $rc -ne 0 ]; then
  echo "NEW ATTENTION KERNEL FAILED -> falling back to the round-1 kernel library for this call"
  export SVI_B200_LIB=$LIBDIR/libsvi_b200_attn_r1.so
else
  for v in "" _attn_poly0 _attn_poly2 _attn_poly4 _attn_poly6 _attn_r1; do
    SVI_B200_LIB=$LIBDIR/libsvi_b200$v.so timeout 120 python tools/gpu_check.py perf_attn_quick 2>&1 | grep PERF
  done | tee gpurun_out/r02_c2_attn_variants.log
fi
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r02_c2_pytest.log 2>&1
echo "pytest exit=$?"
grep -E "inside=|passed|failed|Error|BAD|levels" gpurun_out/r02_c2_pytest.log | tail -n 70
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c2_smoke.log 2>&1
echo "smoke exit=$?"; tail -n 3 gpurun_out/r02_c2_smoke.log
timeout 600 python bench.py > gpurun_out/r02_c2_bench.json 2> gpurun_out/r02_c2_bench.err
echo "bench exit=$?"; head -c 4000 gpurun_out/r02_c2_bench.json; tail -n 5 gpurun_out/r02_c2_bench.err
timeout 400 python bench.py --breakdown --no-cpu-baseline --no-vae > gpurun_out/r02_c2_bench_bd.json 2> gpurun_out/r02_c2_breakdown.txt
echo "breakdown exit=$?"; tail -n 32 gpurun_out/r02_c2_breakdown.txt
timeout 300 python tools/gpu_check.py perf_gemm_epi perf_ew > gpurun_out/r02_c2_perf.log 2>&1
echo "perf exit=$?"; grep PERF gpurun_out/r02_c2_perf.log
