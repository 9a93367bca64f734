#!/bin/bash
# A/B of pair-kernel features on the full step (bench.py without the CPU leg / roofline pass)
mkdir -p gpurun_out
for cfg in "0 1" "1 0" "0 0"; do
  set -- $cfg
  echo "=== B200_TMA_EPI=$1 B200_PREFETCH_B=$2"
  B200_TMA_EPI=$1 B200_PREFETCH_B=$2 timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-roofline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"
done
