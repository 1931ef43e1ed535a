#!/bin/bash
# round-2, 2 GPUs, final code: warp-cooperative partial sums in the q / k norm kernel (one-GPU kernel checks), the plan against the CPU
# oracle, sharded VAE (pair convolution kernel on row bands) == single GPU, bench at N = 2 with sp_parity
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 200 python tools/gpu_check.py abi3 ew > gpurun_out/r02_n2f_check.log 2>&1
echo "check exit=$?"; grep -c "OK " gpurun_out/r02_n2f_check.log; grep -E "BAD|rror" gpurun_out/r02_n2f_check.log | head
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/sp_check.py > gpurun_out/r02_n2f_sp_check.log 2>&1
echo "sp_check exit=$?"; grep -E "oracle|SP_CHECK|BAD|Error" gpurun_out/r02_n2f_sp_check.log | tail -n 20
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 tools/vae_shard_check.py --bench > gpurun_out/r02_n2f_vae_shard.log 2>&1
echo "vae_shard exit=$?"; grep -E "case|VAE_SHARD|Error" gpurun_out/r02_n2f_vae_shard.log | tail -n 8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --steps 3 --warmup 3 --no-vae --no-cpu-baseline > gpurun_out/r02_n2f_bench.json 2> gpurun_out/r02_n2f_bench.err
echo "bench n2 exit=$?"; head -c 1800 gpurun_out/r02_n2f_bench.json; echo; grep sp_parity gpurun_out/r02_n2f_bench.err
