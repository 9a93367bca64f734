#!/bin/bash
# Final evidence of the round: tests, bench, launch list of one step, full ncu captures of the dominant kernels.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-400 gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --skip-cpu --skip-roofline --no-graph --profile-step > gpurun_out/bench_under_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv gpurun_out/launches_summary.json > gpurun_out/launches_summary.txt 2>&1
head -30 gpurun_out/launches_summary.txt; gzip -f gpurun_out/launches.csv
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 3 -c 1 -o gpurun_out/gemm2_lora_full \
    python scripts/one_gemm.py 2048 1280 1280 lora > gpurun_out/ncu_full.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 3 -c 1 -o gpurun_out/gemm2_ff_full \
    python scripts/one_gemm.py 2048 10240 1280 plain >> gpurun_out/ncu_full.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:flash_fwd_kernel -s 2 -c 1 -o gpurun_out/flash_fwd_full \
    python scripts/one_flash.py 4096 >> gpurun_out/ncu_full.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:flash_bwd_kernel -s 2 -c 1 -o gpurun_out/flash_bwd_full \
    python scripts/one_flash.py 4096 >> gpurun_out/ncu_full.log 2>&1
grep -a "Report\|error" gpurun_out/ncu_full.log
