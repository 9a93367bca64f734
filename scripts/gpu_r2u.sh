#!/bin/bash
# tail-split attention backward, parallel token-attention loss kernels, one-hot TI rows: tests, A/B timings, default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_flash_gpu.py tests/test_losses_gpu.py -x -q -m gpu 2>&1 | tail -5
for L in "1024 20" "4096 10"; do
  for T in 0 1; do TIME=1 B200_FLASH_TAILSPLIT=$T timeout 300 python scripts/one_flash.py $L 2>&1 | tail -1; done
done
timeout 900 python bench.py --steps 20 --warmup 3 --skip-gpu-baseline > gpurun_out/bench_r2u.json 2> gpurun_out/bench_r2u.err
echo "bench exit $?"; cut -c1-400 gpurun_out/bench_r2u.json
B200_FLASH_TAILSPLIT=0 timeout 900 python bench.py --steps 20 --warmup 3 --skip-gpu-baseline --skip-roofline --skip-cpu > gpurun_out/bench_r2u_notail.json 2> gpurun_out/bench_r2u_notail.err
echo "bench (no tail split) exit $?"; cut -c1-300 gpurun_out/bench_r2u_notail.json
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
