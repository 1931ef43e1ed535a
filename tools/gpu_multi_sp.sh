#!/bin/bash
# usage: bash tools/gpu_multi_sp.sh <N>   (under gpurun --gpus N): sequence-parallel check + A/B of the K|V exchange (sp = N)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/sp_check.py \
    > gpurun_out/sp_check_$N.log 2>&1
echo "sp_check exit=$?"; grep -E "OK|BAD|SP_CHECK|Error|error|timeout" gpurun_out/sp_check_$N.log | tail -40
for mode in peer nccl; do
  SVI_SP_EXCHANGE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 \
      bench.py --gpus $N --steps 3 --warmup 3 --no-cfg-parallel > gpurun_out/bench_sp${N}_$mode.json 2> gpurun_out/bench_sp${N}_$mode.err
  echo "bench sp$N $mode exit=$?"; tail -n 3 gpurun_out/bench_sp${N}_$mode.err; cat gpurun_out/bench_sp${N}_$mode.json
done
