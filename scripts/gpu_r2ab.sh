#!/bin/bash
# where does the exponential pass of the forward kernel spend its time?  POLY: 0 all MUFU, 4 a quarter on FMA, 1 all on FMA, -1 no exponential
for H in 9 20; do for P in 0 4 1 -1; do
  echo "== H=$H POLY=$P"; TIME=1 B200_FLASH_POLY=$P timeout 300 python scripts/one_flash.py 1024 $H 2>&1 | tail -1
  B200_FLASH_POLY=$P B200_FLASH_TIMELINE=1 timeout 300 python scripts/one_flash.py 1024 $H 2>&1 | grep -A8 "flash_fwd timeline" | tail -5 | cut -c1-100
done; done
