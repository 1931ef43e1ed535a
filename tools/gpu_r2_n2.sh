#!/bin/bash
# round-2, 2 GPUs: the plan against single-GPU kernels and the CPU oracle (incl. the benchmark shape), sharded VAE == single GPU,
# bench at N = 2 with sp_parity, pure sequence-parallel (sp 2) bench for the K/V exchange path
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/sp_check.py > gpurun_out/r02_n2_sp_check.log 2>&1
echo "sp_check exit=$?"; grep -E "oracle|SP_CHECK|BAD|Error" gpurun_out/r02_n2_sp_check.log | tail -n 20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 tools/vae_shard_check.py --bench > gpurun_out/r02_n2_vae_shard.log 2>&1
echo "vae_shard exit=$?"; grep -E "case|VAE_SHARD|Error" gpurun_out/r02_n2_vae_shard.log | tail -n 8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_n2_bench.json 2> gpurun_out/r02_n2_bench.err
echo "bench n2 exit=$?"; head -c 1800 gpurun_out/r02_n2_bench.json; echo; grep sp_parity gpurun_out/r02_n2_bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 2 --steps 3 --warmup 3 --no-cfg-parallel --no-e2e --breakdown > gpurun_out/r02_n2_bench_sp2.json 2> gpurun_out/r02_n2_bench_sp2.err
echo "bench sp2 exit=$?"; head -c 1200 gpurun_out/r02_n2_bench_sp2.json; echo; grep -E "sp_parity|breakdown|ms " gpurun_out/r02_n2_bench_sp2.err | head -n 16
