#!/bin/bash
# production build: flash parity + timings, full GPU suite, default bench; then the timeline build
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_flash_gpu.py -x -q -m gpu 2>&1 | tail -2
TIME=1 timeout 300 python scripts/one_flash.py 1024 20 2>&1 | tail -1
TIME=1 timeout 300 python scripts/one_flash.py 4096 10 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 3 --skip-gpu-baseline > gpurun_out/bench_r2ag.json 2> gpurun_out/bench_r2ag.err
echo "bench exit $?"; cut -c1-260 gpurun_out/bench_r2ag.json; grep -o '"step_loss_delta[^}]*}' gpurun_out/bench_r2ag.json | cut -c1-200
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 bash scripts/flash_timeline.sh > gpurun_out/flash_timelines.txt 2>&1; echo "timeline exit $?"; tail -5 gpurun_out/flash_timelines.txt | cut -c1-150
