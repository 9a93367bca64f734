#!/bin/bash
# packed fp32x2 arithmetic in the attention softmax loops: parity, timelines, timings
timeout 600 python -m pytest tests/test_flash_gpu.py -x -q -m gpu 2>&1 | tail -2
B200_FLASH_TIMELINE=1 timeout 300 python scripts/one_flash.py 1024 9 2>&1 | grep -A8 "flash_bwd timeline" | tail -4 | cut -c1-170
B200_FLASH_TIMELINE=1 timeout 300 python scripts/one_flash.py 1024 20 2>&1 | grep -A8 "flash_fwd timeline" | tail -4 | cut -c1-110
for P in 4 0; do TIME=1 B200_FLASH_POLY=$P timeout 300 python scripts/one_flash.py 1024 20 2>&1 | tail -1 | sed "s/\$/ POLY=$P/"; done
for P in 4 0; do TIME=1 B200_FLASH_POLY=$P timeout 300 python scripts/one_flash.py 4096 10 2>&1 | tail -1 | sed "s/\$/ POLY=$P/"; done
