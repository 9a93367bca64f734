#!/bin/bash
# HBM-bound kernels: CUDA-event timing with L2 flush + ncu --set full captures
mkdir -p gpurun_out
timeout 300 python scripts/one_norms.py > gpurun_out/hbm_kernels.txt 2>&1; echo "one_norms exit $?"; cat gpurun_out/hbm_kernels.txt
for k in layernorm_kernel gn_reduce_kernel gn_bwd_apply_kernel gn_apply_kernel geglu_bwd_kernel geglu_fwd_kernel adamw_kernel lora_wgrad_batch_kernel; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 2 -o gpurun_out/ncu_$k python scripts/one_norms.py > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k exit $?"
done
ls -la gpurun_out/*.ncu-rep | tail -12
