#!/bin/bash
# round-2, 8 GPUs: BASELINE configs[2] (14B-I2V 720p, cfg2 x sp4), configs[1] at N = 8 with a breakdown, configs[3] reduced
# (2 chained 14B clips through the pipeline), configs[4] VAE sweep on row bands
mkdir -p gpurun_out
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
run 900 29711 bench.py --gpus 8 --steps 2 --warmup 3 --workload cfg3 --no-vae > gpurun_out/r02_n8_cfg3.json 2> gpurun_out/r02_n8_cfg3.err
echo "cfg3 exit=$?"; head -c 2500 gpurun_out/r02_n8_cfg3.json; echo; grep -E "sp_parity|Error" gpurun_out/r02_n8_cfg3.err | tail -n 3
run 400 29721 bench.py --gpus 8 --steps 3 --warmup 3 --breakdown > gpurun_out/r02_n8_cfg2.json 2> gpurun_out/r02_n8_cfg2.err
echo "cfg2 n8 exit=$?"; head -c 1500 gpurun_out/r02_n8_cfg2.json; echo; grep -E "sp_parity|breakdown|ms " gpurun_out/r02_n8_cfg2.err | head -n 22
SVI_VAE_SWEEP=17,81,161 run 400 29731 bench.py --gpus 8 --workload cfg5 > gpurun_out/r02_n8_cfg5.json 2> gpurun_out/r02_n8_cfg5.err
echo "cfg5 n8 exit=$?"; head -c 2500 gpurun_out/r02_n8_cfg5.json; echo; tail -n 3 gpurun_out/r02_n8_cfg5.err
run 600 29741 bench.py --gpus 8 --workload cfg4 --clips 2 > gpurun_out/r02_n8_cfg4.json 2> gpurun_out/r02_n8_cfg4.err
echo "cfg4 n8 exit=$?"; head -c 1500 gpurun_out/r02_n8_cfg4.json; echo; tail -n 3 gpurun_out/r02_n8_cfg4.err
