#!/bin/bash
# round-2 call 11: A/B of the pair convolution kernel (make conv_variants): uniform MMA operands x epilogue variant
mkdir -p gpurun_out
LIBDIR=$PWD/stable-video-infinity_b200/lib
for v in "" _conv_u0e1 _conv_u1e0 _conv_u0e0; do
  echo "=== lib${v:-_default(u1e1)}"
  SVI_B200_LIB=$LIBDIR/libsvi_b200$v.so timeout 200 python tools/gpu_check.py perf_conv 2>&1 | grep PERF | sed 's/  v1 [0-9]* us = [0-9]* TF\/s//'
  SVI_B200_LIB=$LIBDIR/libsvi_b200$v.so timeout 200 python tools/vae_bench.py --iters 2 2>/dev/null | tail -c 420 | grep -o '"decode": {"ms": [0-9.]*\|"encode": {"ms": [0-9.]*'
done 2>&1 | tee gpurun_out/r02_c11_conv_variants.log
