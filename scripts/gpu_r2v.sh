#!/bin/bash
# tail-split sweep of the attention backward (B200_FLASH_TAILSPLIT = most pieces per item of the partial round)
timeout 600 python -m pytest tests/test_flash_gpu.py -x -q -m gpu 2>&1 | tail -2
for T in 1 2 3 4 6 8; do TIME=1 B200_FLASH_TAILSPLIT=$T timeout 300 python scripts/one_flash.py 1024 20 2>&1 | tail -1; done
for T in 1 2 3; do TIME=1 B200_FLASH_TAILSPLIT=$T timeout 300 python scripts/one_flash.py 4096 10 2>&1 | tail -1; done
