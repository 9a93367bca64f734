#!/bin/bash
# sanitizers over the kernels changed late in round 2: TS-mode attention forward, backward (3-stage ring, tail split, TS-mode dQ),
# delta kernel, token-attention loss kernels
mkdir -p gpurun_out
T="tests/test_flash_gpu.py::test_flash_attention_fwd_bwd[2-4-256-256] tests/test_flash_gpu.py::test_flash_attention_fwd_bwd[2-3-384-77] tests/test_flash_gpu.py::test_flash_attention_fwd_bwd[2-2-200-150] tests/test_flash_gpu.py::test_flash_attention_fwd_bwd[1-1-128-65] tests/test_flash_gpu.py::test_flash_attention_fwd_bwd[2-20-512-512] tests/test_flash_gpu.py::test_flash_forward_rescales_when_later_keys_dominate tests/test_losses_gpu.py"
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $T -q -m gpu > gpurun_out/sanitizer_memcheck_r02p.log 2>&1; echo "memcheck exit $?"; grep -a "ERROR SUMMARY\|passed\|failed" gpurun_out/sanitizer_memcheck_r02p.log | tail -3
timeout 400 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest $T -q -m gpu > gpurun_out/sanitizer_synccheck_r02p.log 2>&1; echo "synccheck exit $?"; grep -a "ERROR SUMMARY\|passed\|failed" gpurun_out/sanitizer_synccheck_r02p.log | tail -3
