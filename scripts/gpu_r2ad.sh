#!/bin/bash
# attention backward with a 3-stage Q / dO ring: parity, timeline, timings
timeout 600 python -m pytest tests/test_flash_gpu.py -x -q -m gpu 2>&1 | tail -2
B200_FLASH_TIMELINE=1 timeout 300 python scripts/one_flash.py 1024 9 2>&1 | grep -A8 "flash_bwd timeline" | tail -9
TIME=1 timeout 300 python scripts/one_flash.py 1024 20 2>&1 | tail -1
TIME=1 timeout 300 python scripts/one_flash.py 4096 10 2>&1 | tail -1
