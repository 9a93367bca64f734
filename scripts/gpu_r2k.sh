#!/bin/bash
# round 2, call 11: dense weight gradients on the CTA-pair kernel (MN-major A, fp32 accumulate), AdamW without per-element rcp of
# the bias correction; BASELINE configs 5 / 2 with their CPU legs (step_loss_delta), config 1 (CPU reference leg)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm2_gpu.py tests/test_kernels_gpu.py tests/test_z3_dense_gpu.py -q -x > gpurun_out/pytest_new.log 2>&1; echo "pytest new exit $?"; tail -5 gpurun_out/pytest_new.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --full-ft --batch 1 --steps 5 --warmup 3 --skip-gpu-baseline --no-graph --delta-timeout 700 > gpurun_out/bench_full_ft.json 2> gpurun_out/bench_full_ft.err
echo "bench full-ft exit $?"; cut -c1-200 gpurun_out/bench_full_ft.json; grep -o '"step_loss_delta": {[^}]*}' gpurun_out/bench_full_ft.json; grep -o '"cpu_baseline": {[^}]*}' gpurun_out/bench_full_ft.json | cut -c1-200
B200_PDL=0 timeout 400 python scripts/profile_step.py --full-ft --batch 1 --tag sdxl_ft_nopdl > gpurun_out/profile_ft.log 2>&1; echo "profile ft exit $?"; head -16 gpurun_out/step_kernels_sdxl_ft_nopdl.txt
timeout 600 python bench.py --family sd15 --res 512 --batch 4 --rank 16 --steps 10 --warmup 3 > gpurun_out/bench_config2_sd15.json 2> gpurun_out/bench_config2_sd15.err
echo "bench config 2 exit $?"; cut -c1-200 gpurun_out/bench_config2_sd15.json; grep -o '"step_loss_delta": {[^}]*}' gpurun_out/bench_config2_sd15.json; grep -o '"gpu_baseline": {[^}]*}' gpurun_out/bench_config2_sd15.json | cut -c1-160
timeout 600 python bench.py --impl reference --family sd15 --rank 4 --res 512 --steps 5 --warmup 1 > gpurun_out/bench_config1_cpu_reference.json 2> gpurun_out/bench_config1_cpu_reference.err
echo "bench config 1 (CPU reference leg) exit $?"; cut -c1-300 gpurun_out/bench_config1_cpu_reference.json
timeout 600 python bench.py --family sd15 --res 512 --batch 1 --rank 4 --steps 10 --warmup 3 --skip-gpu-baseline --skip-roofline > gpurun_out/bench_config1_gpu.json 2> gpurun_out/bench_config1_gpu.err
echo "bench config 1 shape on the GPU exit $?"; cut -c1-200 gpurun_out/bench_config1_gpu.json; grep -o '"step_loss_delta": {[^}]*}' gpurun_out/bench_config1_gpu.json
