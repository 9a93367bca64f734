#!/bin/bash
# per-block timeline of the attention backward kernel (CTA 0), L = 1024: one round (H=9)
B200_FLASH_TIMELINE=1 timeout 300 python scripts/one_flash.py 1024 9 2>&1 | grep -A8 "flash_bwd timeline" | tail -9
TIME=1 timeout 300 python scripts/one_flash.py 1024 20 2>&1 | tail -1
TIME=1 timeout 300 python scripts/one_flash.py 4096 10 2>&1 | tail -1
