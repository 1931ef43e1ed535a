#!/bin/bash
# round-2 call 6: CTA-pair convolution kernel with the input window reused across the horizontal taps (bring-up of the row-shifted
# operand descriptors, parity, timing against the single-CTA kernel at the VAE's layer shapes, VAE bench), f3 hand-off kernels
mkdir -p gpurun_out
SVI_PROBE_FLAGS=0 timeout 200 python tools/gpu_check.py conv_pair_probe > gpurun_out/r02_c6_pair_probe.log 2>&1
SVI_PROBE_FLAGS=1 timeout 200 python tools/gpu_check.py conv_pair_probe >> gpurun_out/r02_c6_pair_probe.log 2>&1
echo "probe exit=$?"; grep -E "^---|BAD|raised|rror" gpurun_out/r02_c6_pair_probe.log | head -40
F0=$(grep -c "PROBE flags=0 allok=1" gpurun_out/r02_c6_pair_probe.log)
F1=$(grep -c "PROBE flags=1 allok=1" gpurun_out/r02_c6_pair_probe.log)
echo "flags0 all ok: $F0   flags1 all ok: $F1"
if [ "$F0" = "0" ] && [ "$F1" != "0" ]; then export SVI_CONV_FLAGS=1; fi
if [ "$F0" = "0" ] && [ "$F1" = "0" ]; then export SVI_CONV_VARIANT=1; echo "pair kernel failed both ways: rest of the call on the single-CTA kernel"; fi
echo "SVI_CONV_FLAGS=$SVI_CONV_FLAGS SVI_CONV_VARIANT=$SVI_CONV_VARIANT"
timeout 300 python tools/gpu_check.py perf_conv > gpurun_out/r02_c6_perf_conv.log 2>&1
echo "perf_conv exit=$?"; grep PERF gpurun_out/r02_c6_perf_conv.log
timeout 300 python tools/gpu_check.py conv abi3 > gpurun_out/r02_c6_check.log 2>&1
echo "check exit=$?"; grep -c "OK " gpurun_out/r02_c6_check.log; grep -E "BAD|rror" gpurun_out/r02_c6_check.log | head
timeout 600 python -m pytest tests/test_vae_gpu.py tests/test_pipeline_gpu.py tests/test_harness_gpu.py tests/test_dance_gpu.py -m gpu -q -s > gpurun_out/r02_c6_pytest.log 2>&1
echo "pytest exit=$?"; grep -E "passed|failed|Error|BAD|levels" gpurun_out/r02_c6_pytest.log | tail -n 20
SVI_CONV_VARIANT=1 timeout 300 python tools/vae_bench.py --iters 2 > gpurun_out/r02_c6_vae_bench_v1.json 2> gpurun_out/r02_c6_vae_bench_v1.err
echo "vae bench v1 exit=$?"; tail -c 1500 gpurun_out/r02_c6_vae_bench_v1.json; echo
timeout 300 python tools/vae_bench.py --iters 2 > gpurun_out/r02_c6_vae_bench.json 2> gpurun_out/r02_c6_vae_bench.err
echo "vae bench auto exit=$?"; tail -c 1500 gpurun_out/r02_c6_vae_bench.json; echo
