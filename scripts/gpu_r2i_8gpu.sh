#!/bin/bash
# 8 x B200: BASELINE config 4 (SDXL LoRA r=32, global batch 16 = 2 per GPU, LoRA-grad all-reduce inside the step's graph)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --rank 32 --steps 10 --warmup 3 --skip-roofline > gpurun_out/bench_8gpu_r32.json 2> gpurun_out/bench_8gpu_r32.err
echo "bench 8gpu r32 (config 4) exit $?"; cut -c1-300 gpurun_out/bench_8gpu_r32.json; grep "process group" gpurun_out/bench_8gpu_r32.err | head -3
