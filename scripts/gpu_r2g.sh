#!/bin/bash
# round 2, call 7: loss kernels (token-attention regulariser + gradient map, token-std regulariser)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_losses_gpu.py -q -x > gpurun_out/pytest_new.log 2>&1; echo "pytest losses exit $?"; tail -12 gpurun_out/pytest_new.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-gpu-baseline --skip-roofline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default exit $?"; cut -c1-200 gpurun_out/bench_default.json
B200_TAL_KERNEL=0 timeout 400 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-gpu-baseline --skip-roofline > gpurun_out/bench_no_tal_kernel.json 2> gpurun_out/bench_no_tal_kernel.err
echo "bench B200_TAL_KERNEL=0 exit $?"; cut -c1-200 gpurun_out/bench_no_tal_kernel.json
