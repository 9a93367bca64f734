#!/bin/bash
# delta kernel with 8 lanes per (row, head) that also zeroes the dQ accumulator: parity, timing, default bench
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_flash_gpu.py -x -q -m gpu 2>&1 | tail -3
TIME=1 timeout 60 python scripts/one_flash.py 1024 20 2>&1 | tail -1
TIME=1 timeout 60 python scripts/one_flash.py 4096 10 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 3 --skip-gpu-baseline --skip-roofline > gpurun_out/bench_r2al.json 2> gpurun_out/bench_r2al.err
echo "bench exit $?"; cut -c1-260 gpurun_out/bench_r2al.json; grep -o '"step_loss_delta[^}]*}' gpurun_out/bench_r2al.json | cut -c1-160
