#!/bin/bash
# per-block timeline of the forward attention kernel (CTA 0): one resident CTA per SM (H=9) and two (H=20)
for H in 9 20; do B200_FLASH_TIMELINE=1 timeout 300 python scripts/one_flash.py 1024 $H 2>&1 | grep -A8 "flash_fwd timeline" | tail -9; done
