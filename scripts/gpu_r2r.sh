#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gemm2_gpu.py tests/test_kernels_gpu.py -q -x > gpurun_out/pytest_new.log 2>&1; echo "pytest new exit $?"; tail -4 gpurun_out/pytest_new.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --skip-cpu --skip-gpu-baseline --skip-roofline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default exit $?"; cut -c1-200 gpurun_out/bench_default.json
B200_FUSE_GEGLU_FWD=0 timeout 600 python bench.py --steps 20 --warmup 3 --skip-cpu --skip-gpu-baseline --skip-roofline > gpurun_out/bench_no_fuse_geglu_fwd.json 2> gpurun_out/bench_no_fuse_geglu_fwd.err
echo "bench B200_FUSE_GEGLU_FWD=0 exit $?"; cut -c1-200 gpurun_out/bench_no_fuse_geglu_fwd.json
