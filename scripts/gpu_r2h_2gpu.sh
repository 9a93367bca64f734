#!/bin/bash
# 2 x B200: data-parallel bench (in-graph all-reduce), clean process-group teardown, rank-32 shard
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "bench 2gpu exit $?"; cut -c1-300 gpurun_out/bench_2gpu.json; grep "process group" gpurun_out/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --rank 32 --steps 10 --warmup 3 --skip-roofline > gpurun_out/bench_2gpu_r32.json 2> gpurun_out/bench_2gpu_r32.err
echo "bench 2gpu r32 exit $?"; cut -c1-300 gpurun_out/bench_2gpu_r32.json; grep "process group" gpurun_out/bench_2gpu_r32.err
