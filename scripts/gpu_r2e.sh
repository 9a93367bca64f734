#!/bin/bash
# round 2, call 5: fused q|k|v projection (wide side path), flash row strides, B200_PDL=0 profile
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm2_gpu.py tests/test_flash_gpu.py tests/test_z6_shapes_gpu.py -q -x > gpurun_out/pytest_new.log 2>&1; echo "pytest new exit $?"; tail -6 gpurun_out/pytest_new.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-gpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default exit $?"; cut -c1-200 gpurun_out/bench_default.json; cp gpurun_out/gemm_by_shape.json gpurun_out/gemm_by_shape_sdxl_r16.json
B200_FUSE_QKV=0 timeout 400 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-gpu-baseline --skip-roofline > gpurun_out/bench_no_fuse_qkv.json 2> gpurun_out/bench_no_fuse_qkv.err
echo "bench B200_FUSE_QKV=0 exit $?"; cut -c1-200 gpurun_out/bench_no_fuse_qkv.json
B200_PDL=0 timeout 400 python scripts/profile_step.py --tag sdxl_r16_nopdl > gpurun_out/profile_step.log 2>&1; echo "profile exit $?"; head -50 gpurun_out/step_kernels_sdxl_r16_nopdl.txt
