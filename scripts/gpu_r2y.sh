#!/bin/bash
# forward attention with P in tensor memory (TS-mode MMA): parity tests, A/B timing, default bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_flash_gpu.py -x -q -m gpu 2>&1 | tail -4
for L in "1024 20" "4096 10" "1024 9"; do
  for T in 0 1; do TIME=1 B200_FLASH_FWD_TS=$T timeout 300 python scripts/one_flash.py $L 2>&1 | tail -1 | sed "s/\$/ FWD_TS=$T/"; done
done
timeout 900 python bench.py --steps 20 --warmup 3 --skip-gpu-baseline --skip-roofline > gpurun_out/bench_r2y.json 2> gpurun_out/bench_r2y.err
echo "bench exit $?"; cut -c1-300 gpurun_out/bench_r2y.json; grep -o '"step_loss_delta[^}]*}' gpurun_out/bench_r2y.json | head -2
