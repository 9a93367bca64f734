#!/bin/bash
mkdir -p gpurun_out
B200_STREAMK=0 timeout 300 python scripts/determinism_trace.py sdxl > gpurun_out/determinism_trace.txt 2>&1
B200_STREAMK=1 timeout 300 python scripts/determinism_trace.py sdxl >> gpurun_out/determinism_trace.txt 2>&1
B200_STREAMK=0 timeout 300 python scripts/determinism_trace.py sd15 64 >> gpurun_out/determinism_trace.txt 2>&1
grep -v "Warning\|warn" gpurun_out/determinism_trace.txt | tail -30
