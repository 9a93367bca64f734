#!/bin/bash
# One gpurun call: GPU parity tests, the headline bench, the ncu launch list of one step, a full ncu capture of the
# LoRA-fused GEMM, and the GEMM micro-probes.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/nvidia_smi.txt 2>&1
nproc >> gpurun_out/nvidia_smi.txt

if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  tail -5 gpurun_out/pytest_gpu.log
fi

timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; cat gpurun_out/bench.json

if [ "${SKIP_NCU:-0}" != "1" ]; then
  # every launch of one eager step (3 warm-up steps skipped by the summariser: it keeps the last complete step)
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 3 --skip-cpu --skip-roofline --no-graph > gpurun_out/bench_under_ncu.log 2>&1
  python scripts/summarize_launches.py gpurun_out/launches.csv gpurun_out/launches_summary.json > gpurun_out/launches_summary.txt 2>&1
  head -40 gpurun_out/launches_summary.txt
  # keep only the summarised step's rows (the raw list of 5 steps is large)
  python - <<'EOF'
import gzip, shutil
with open('gpurun_out/launches.csv','rb') as f, gzip.open('gpurun_out/launches.csv.gz','wb') as g:
    shutil.copyfileobj(f, g)
EOF
  rm -f gpurun_out/launches.csv
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 3 -c 2 \
      -o gpurun_out/gemm_lora_full python scripts/one_gemm.py 2048 1280 1280 lora > gpurun_out/ncu_full.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 3 -c 1 \
      -o gpurun_out/gemm_ff_full python scripts/one_gemm.py 2048 10240 1280 plain >> gpurun_out/ncu_full.log 2>&1
fi

timeout 300 python scripts/bench_gemm.py > gpurun_out/bench_gemm.log 2>&1
cat gpurun_out/bench_gemm.log
timeout 300 python scripts/probe_mma.py > gpurun_out/probe_mma.log 2>&1
cat gpurun_out/probe_mma.log
timeout 300 python scripts/probe_stamps.py > gpurun_out/probe_stamps.log 2>&1
cat gpurun_out/probe_stamps.log
