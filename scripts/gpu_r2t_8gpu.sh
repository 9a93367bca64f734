#!/bin/bash
# 8 x B200 on the final tree: headline config (r=16, B=2 per GPU) and BASELINE config 4 (r=32, global batch 16)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 --skip-roofline > gpurun_out/bench_8gpu_r16.json 2> gpurun_out/bench_8gpu_r16.err
echo "bench 8gpu r16 exit $?"; cut -c1-260 gpurun_out/bench_8gpu_r16.json; grep -c "destroyed cleanly" gpurun_out/bench_8gpu_r16.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --rank 32 --steps 20 --warmup 3 --skip-roofline > gpurun_out/bench_8gpu_r32.json 2> gpurun_out/bench_8gpu_r32.err
echo "bench 8gpu r32 (config 4) exit $?"; cut -c1-260 gpurun_out/bench_8gpu_r32.json; grep -c "destroyed cleanly" gpurun_out/bench_8gpu_r32.err
