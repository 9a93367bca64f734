#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_flash_gpu.py -q -x > gpurun_out/pytest_new.log 2>&1; echo "pytest flash exit $?"; tail -4 gpurun_out/pytest_new.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-gpu-baseline --skip-roofline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default exit $?"; cut -c1-200 gpurun_out/bench_default.json
timeout 300 python bench.py --family sd15 --res 512 --batch 4 --rank 16 --steps 10 --warmup 3 --skip-cpu --skip-gpu-baseline --skip-roofline > gpurun_out/bench_config2_sd15.json 2> gpurun_out/bench_config2_sd15.err
echo "bench config 2 exit $?"; cut -c1-200 gpurun_out/bench_config2_sd15.json
