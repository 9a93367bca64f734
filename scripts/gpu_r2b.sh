#!/bin/bash
# round 2, call 2: vectorised optimizers, head re-pitch + narrow-head fused attention, determinism probe, in-situ profile,
# config 2 with fused attention, default bench with the gpu_baseline leg
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -4 gpurun_out/pytest_gpu.log
B200_STREAMK=0 timeout 200 python scripts/determinism.py > gpurun_out/determinism.txt 2>&1
B200_STREAMK=1 timeout 200 python scripts/determinism.py >> gpurun_out/determinism.txt 2>&1
grep -v Warning gpurun_out/determinism.txt | tail -12
timeout 400 python scripts/profile_step.py --tag sdxl_r16 > gpurun_out/profile_step.log 2>&1; echo "profile exit $?"; head -45 gpurun_out/step_kernels_sdxl_r16.txt
timeout 300 python bench.py --family sd15 --res 512 --batch 4 --rank 16 --steps 10 --warmup 3 --skip-cpu --skip-roofline --skip-gpu-baseline > gpurun_out/bench_config2_sd15.json 2> gpurun_out/bench_config2_sd15.err
echo "bench config 2 exit $?"; cut -c1-200 gpurun_out/bench_config2_sd15.json
timeout 900 python bench.py --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default exit $?"; cut -c1-200 gpurun_out/bench_default.json; grep -o '"gpu_baseline": {[^}]*}' gpurun_out/bench_default.json
