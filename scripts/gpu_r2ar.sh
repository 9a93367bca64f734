#!/bin/bash
# final committed tree: attention tests (incl. the three large Lk = 77 shapes) on the general kernels
timeout 150 python -m pytest tests/test_flash_gpu.py -x -q -m gpu 2>&1 | tail -3
