#!/bin/bash
# ncu --set full of the attention kernels at their best-filled shape (B=2, H=10, L=Lk=4096: 640 CTAs)
mkdir -p gpurun_out
timeout 150 ncu --set full --clock-control none --import-source on -k regex:flash_fwd -s 2 -c 1 -o gpurun_out/final_flash_fwd_l4096 python scripts/one_flash.py 4096 10 > gpurun_out/ncu_l4096.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:flash_bwd_kernel -s 2 -c 1 -o gpurun_out/final_flash_bwd_l4096 python scripts/one_flash.py 4096 10 >> gpurun_out/ncu_l4096.log 2>&1
grep -a "Report\|rror" gpurun_out/ncu_l4096.log | head -4
