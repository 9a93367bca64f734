#!/bin/bash
# attention backward with dS as a TMEM-resident A operand of dQ = dS.K (TS-mode MMA)
timeout 200 python -m pytest tests/test_flash_gpu.py -x -q -m gpu 2>&1 | tail -3
TIME=1 timeout 60 python scripts/one_flash.py 1024 20 2>&1 | tail -1
TIME=1 timeout 60 python scripts/one_flash.py 4096 10 2>&1 | tail -1
