#!/bin/bash
# round-2 call 3: attention poly-share / register sweep; conv epilogue fusion + LN fold correctness and A/B; VAE timing;
# context_state breakdown; ncu capture of the attention kernel
mkdir -p gpurun_out
LIBDIR=$PWD/stable-video-infinity_b200/lib
for v in "" _attn_poly4 _attn_poly8 _attn_poly10 _attn_poly12 _attn_poly16 _attn_regs232 _attn_regs208 _attn_r1; do
  SVI_B200_LIB=$LIBDIR/libsvi_b200$v.so timeout 120 python tools/gpu_check.py perf_attn_quick 2>&1 | grep PERF
done | tee gpurun_out/r02_c3_attn_variants.log
timeout 300 python tools/gpu_check.py conv abi3 attn_bench gemm_epi > gpurun_out/r02_c3_check.log 2>&1
echo "check exit=$?"; grep -c "OK " gpurun_out/r02_c3_check.log; grep -E "BAD|Error|error" gpurun_out/r02_c3_check.log | head -20
timeout 900 python -m pytest tests/test_dit_gpu.py tests/test_vae_gpu.py tests/test_pipeline_gpu.py tests/test_kernels_gpu.py -m gpu -q -s > gpurun_out/r02_c3_pytest.log 2>&1
echo "pytest exit=$?"
grep -E "inside=|passed|failed|Error|BAD|levels" gpurun_out/r02_c3_pytest.log | tail -n 40
timeout 400 python bench.py --no-vae > gpurun_out/r02_c3_bench.json 2> gpurun_out/r02_c3_bench.err
echo "bench exit=$?"; head -c 2500 gpurun_out/r02_c3_bench.json; echo; tail -n 3 gpurun_out/r02_c3_bench.err
SVI_LN_FOLD=0 timeout 400 python bench.py --no-vae --no-cpu-baseline --no-e2e > gpurun_out/r02_c3_bench_nofold.json 2> gpurun_out/r02_c3_bench_nofold.err
echo "bench nofold exit=$?"; head -c 700 gpurun_out/r02_c3_bench_nofold.json; echo
timeout 400 python bench.py --breakdown --no-cpu-baseline --no-vae --no-e2e > gpurun_out/r02_c3_bench_bd.json 2> gpurun_out/r02_c3_breakdown.txt
echo "breakdown exit=$?"; tail -n 24 gpurun_out/r02_c3_breakdown.txt
timeout 300 python tools/vae_bench.py > gpurun_out/r02_c3_vae_480p.json 2> gpurun_out/r02_c3_vae.err
echo "vae exit=$?"; cat gpurun_out/r02_c3_vae_480p.json; tail -n 3 gpurun_out/r02_c3_vae.err
timeout 200 python - > gpurun_out/r02_c3_ctx_breakdown.txt 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'stable-video-infinity_b200')
import bench
from tools import synth
dev = torch.device('cuda', 0)
model = bench.build_model(synth.CFG_T2V_1_3B, dev)
eng = model.engine(dev)
ctx = torch.randn(1, 512, 4096).pin_memory()
for _ in range(3):
    eng.context_state(ctx.to(dev))
torch.cuda.synchronize()
import time
t0 = time.perf_counter(); eng.context_state(ctx.to(dev)); torch.cuda.synchronize(); print('wall ms', (time.perf_counter() - t0) * 1e3)
eng.k.events = []
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); eng.context_state(ctx.to(dev)); b.record()
bd = eng.k.breakdown()
print('gpu ms between events', a.elapsed_time(b))
for tag, (ms, n) in sorted(bd.items(), key=lambda kv: -kv[1][0]):
    print(f'{ms:8.3f} ms n={n:3d} {tag}')
PY
cat gpurun_out/r02_c3_ctx_breakdown.txt | tail -n 12
ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 3 -c 1 -o gpurun_out/r02_prof_attn -f \
    python tools/gpu_check.py perf_attn_quick > gpurun_out/r02_c3_ncu_attn_stdout.log 2>&1
echo "ncu attn exit=$?"; ls -la gpurun_out/r02_prof_attn.ncu-rep
