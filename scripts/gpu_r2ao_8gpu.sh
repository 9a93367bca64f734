#!/bin/bash
# final tree: smoke(), then 8 x B200 - headline config (r=16, B=2 per GPU) and BASELINE config 4 (r=32, global batch 16)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 --skip-roofline > gpurun_out/bench_8gpu_r16.json 2> gpurun_out/bench_8gpu_r16.err
echo "bench 8gpu r16 exit $?"; cut -c1-260 gpurun_out/bench_8gpu_r16.json; grep -c "destroyed cleanly" gpurun_out/bench_8gpu_r16.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --rank 32 --steps 20 --warmup 3 --skip-roofline > gpurun_out/bench_8gpu_r32.json 2> gpurun_out/bench_8gpu_r32.err
echo "bench 8gpu r32 (config 4) exit $?"; cut -c1-260 gpurun_out/bench_8gpu_r32.json; grep -c "destroyed cleanly" gpurun_out/bench_8gpu_r32.err
