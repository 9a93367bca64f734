#!/bin/bash
# Final evidence of round 2 (profiles/r02n_*, earlier passes r02g_ / r02k_): GPU suite, default bench with every leg, reference arm, launch list of one step,
# ncu --set full of the dominant kernels of the FINAL build.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_default.json
cp gpurun_out/gemm_by_shape.json gpurun_out/gemm_by_shape_sdxl_r16.json
grep -o '"step_loss_delta": {[^}]*}' gpurun_out/bench_default.json; grep -o '"gpu_baseline": {[^}]*}' gpurun_out/bench_default.json | cut -c1-200; grep -o '"roofline": {[^}]*}' gpurun_out/bench_default.json | cut -c1-300
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err; echo "reference arm exit $?"; cut -c1-250 gpurun_out/bench_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --skip-cpu --skip-roofline --skip-gpu-baseline --no-graph --profile-step > gpurun_out/bench_under_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv gpurun_out/launches_summary.json > gpurun_out/launches_summary.txt 2>&1
head -24 gpurun_out/launches_summary.txt; gzip -f gpurun_out/launches.csv
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 3 -c 1 -o gpurun_out/final_gemm2_lora \
    python scripts/one_gemm.py 2048 1280 1280 lora 16 > gpurun_out/ncu_full.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 3 -c 1 -o gpurun_out/final_gemm2_qkv \
    python scripts/one_gemm.py 2048 3840 1280 lora 48 >> gpurun_out/ncu_full.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 3 -c 1 -o gpurun_out/final_gemm2_ff \
    python scripts/one_gemm.py 2048 10240 1280 plain >> gpurun_out/ncu_full.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:flash_fwd -s 2 -c 1 -o gpurun_out/final_flash_fwd \
    python scripts/one_flash.py 1024 20 >> gpurun_out/ncu_full.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:flash_bwd_kernel -s 2 -c 1 -o gpurun_out/final_flash_bwd \
    python scripts/one_flash.py 1024 20 >> gpurun_out/ncu_full.log 2>&1
grep -a "Report\|rror" gpurun_out/ncu_full.log | head
