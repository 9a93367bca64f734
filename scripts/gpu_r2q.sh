#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -4 gpurun_out/pytest_gpu.log
B200_STREAMK=0 timeout 300 python scripts/determinism_trace.py sdxl > gpurun_out/determinism_trace.txt 2>&1
B200_STREAMK=1 timeout 300 python scripts/determinism_trace.py sdxl >> gpurun_out/determinism_trace.txt 2>&1
grep -v "Warning\|warn\|first 6\|run_backward" gpurun_out/determinism_trace.txt | tail -8
timeout 600 python bench.py --steps 20 --warmup 3 --skip-cpu --skip-gpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default exit $?"; cut -c1-200 gpurun_out/bench_default.json; cp gpurun_out/gemm_by_shape.json gpurun_out/gemm_by_shape_sdxl_r16.json
B200_PDL=0 timeout 400 python scripts/profile_step.py --tag sdxl_r16_nopdl > gpurun_out/profile_step.log 2>&1; echo "profile exit $?"; head -14 gpurun_out/step_kernels_sdxl_r16_nopdl.txt
