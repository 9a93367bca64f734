#!/bin/bash
# forward attention: share of the exponentials on the FMA pipe (B200_FLASH_POLY = every n-th; 0 = none)
timeout 600 python -m pytest tests/test_flash_gpu.py -x -q -m gpu 2>&1 | tail -2
for L in "1024 20" "4096 10"; do
  for P in 0 4 3 2; do TIME=1 B200_FLASH_POLY=$P timeout 300 python scripts/one_flash.py $L 2>&1 | tail -1 | sed "s/\$/ POLY=$P/"; done
done
B200_FLASH_TIMELINE=1 timeout 300 python scripts/one_flash.py 1024 20 2>&1 | grep -A8 "flash_fwd timeline" | tail -9
