#!/bin/bash
mkdir -p gpurun_out
B200_PDL=0 timeout 400 python scripts/profile_step.py --family sd15 --res 512 --batch 4 --rank 16 --tag sd15_r16_nopdl > gpurun_out/profile_sd15.log 2>&1; echo "profile sd15 exit $?"; head -40 gpurun_out/step_kernels_sd15_r16_nopdl.txt
