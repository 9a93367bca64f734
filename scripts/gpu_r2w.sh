#!/bin/bash
# rounds model of the attention backward at L = 1024: items = 16 x H on 148 SMs
for H in 9 18 19 20 27 28; do for T in 1 8; do TIME=1 B200_FLASH_TAILSPLIT=$T timeout 300 python scripts/one_flash.py 1024 $H 2>&1 | tail -1; done; done
