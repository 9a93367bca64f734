#!/bin/bash
# final tree: the rescale-path attention test, then BASELINE configs 1, 2 and 5 (config 3 = default bench, config 4 = 8 GPUs)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_flash_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 600 python bench.py --family sd15 --res 512 --batch 4 --rank 16 --steps 10 --warmup 3 --skip-gpu-baseline > gpurun_out/bench_config2_sd15.json 2> gpurun_out/bench_config2_sd15.err
echo "config 2 exit $?"; cut -c1-250 gpurun_out/bench_config2_sd15.json; grep -o '"step_loss_delta": {[^}]*}' gpurun_out/bench_config2_sd15.json | cut -c1-160
timeout 600 python bench.py --family sd15 --res 512 --batch 1 --rank 4 --steps 10 --warmup 3 --skip-gpu-baseline --skip-roofline > gpurun_out/bench_config1_gpu.json 2> gpurun_out/bench_config1_gpu.err
echo "config 1 exit $?"; cut -c1-250 gpurun_out/bench_config1_gpu.json; grep -o '"step_loss_delta": {[^}]*}' gpurun_out/bench_config1_gpu.json | cut -c1-160
timeout 900 python bench.py --full-ft --batch 1 --steps 5 --warmup 3 --skip-cpu --skip-gpu-baseline --skip-roofline > gpurun_out/bench_config5_full_ft.json 2> gpurun_out/bench_config5_full_ft.err
echo "config 5 exit $?"; cut -c1-250 gpurun_out/bench_config5_full_ft.json
