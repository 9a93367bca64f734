#!/bin/bash
# last sanity pass on the final tree: smoke() and the attention / loss tests
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 150 python -m pytest tests/test_flash_gpu.py tests/test_losses_gpu.py -x -q -m gpu 2>&1 | tail -2
