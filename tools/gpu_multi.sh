#!/bin/bash
# usage: bash tools/gpu_multi.sh <N>   (under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/sp_check.py \
    > gpurun_out/sp_check_$N.log 2>&1
echo "sp_check exit=$?"; grep -E "OK|BAD|SP_CHECK|Error" gpurun_out/sp_check_$N.log | tail -20
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 3 --warmup 3 \
    > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err
echo "bench N=$N exit=$?"; tail -n 3 gpurun_out/bench_$N.err; cat gpurun_out/bench_$N.json
