#!/bin/bash
# Per-block clock64 timelines of the attention kernels (CTA 0).  Rebuilds the library IN THE CURRENT TREE with
# -DB200_FLASH_TIMELINE_BUILD=1 (run it on a scratch copy, e.g. under gpurun) and prints forward / backward timelines for
# one resident CTA per SM (H=9) and the SDXL level-2 shape (H=20), L = Lk = 1024.
set -e
cd "$(dirname "$0")/.."
C=sd_lora_trainer_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
     -DB200_FLASH_TIMELINE_BUILD=1 -c $C/flash_attn.cu -o $C/flash_attn.o
nvcc -shared -o sd_lora_trainer_b200/libb200_lora.so $C/gemm_host.o $C/flash_attn.o $C/norms.o $C/elementwise.o $C/optim.o $C/wgrad_batch.o $C/losses.o
for H in 9 20; do
  B200_FLASH_TIMELINE=1 python scripts/one_flash.py 1024 $H 2>&1 | grep -A8 "flash_fwd timeline" | tail -9
  B200_FLASH_TIMELINE=1 python scripts/one_flash.py 1024 $H 2>&1 | grep -A8 "flash_bwd timeline" | tail -9
done
