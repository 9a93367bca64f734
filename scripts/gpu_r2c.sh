#!/bin/bash
# round 2, call 3: query-split cross-attention backward, conv-LoRA shift-sum, batched LoRA weight gradients, shared dscores
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_wgrad_batch_gpu.py tests/test_flash_gpu.py -q -x > gpurun_out/pytest_new.log 2>&1; echo "pytest new exit $?"; tail -4 gpurun_out/pytest_new.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-gpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default exit $?"; cut -c1-200 gpurun_out/bench_default.json
B200_WGRAD_BATCH=0 timeout 400 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-gpu-baseline --skip-roofline > gpurun_out/bench_no_wgrad_batch.json 2> gpurun_out/bench_no_wgrad_batch.err
echo "bench B200_WGRAD_BATCH=0 exit $?"; cut -c1-200 gpurun_out/bench_no_wgrad_batch.json
timeout 400 python scripts/profile_step.py --tag sdxl_r16 > gpurun_out/profile_step.log 2>&1; echo "profile exit $?"; head -30 gpurun_out/step_kernels_sdxl_r16.txt
timeout 500 python bench.py --full-ft --batch 1 --steps 5 --warmup 3 --skip-cpu --skip-gpu-baseline --no-graph > gpurun_out/bench_full_ft.json 2> gpurun_out/bench_full_ft.err
echo "bench full-ft exit $?"; cut -c1-200 gpurun_out/bench_full_ft.json; tail -3 gpurun_out/bench_full_ft.err
timeout 400 python scripts/profile_step.py --full-ft --batch 1 --tag sdxl_ft > gpurun_out/profile_ft.log 2>&1; echo "profile ft exit $?"; head -30 gpurun_out/step_kernels_sdxl_ft.txt
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_wgrad_batch_gpu.py "tests/test_flash_gpu.py::test_flash_attention_fwd_bwd[2-3-384-77]" "tests/test_flash_gpu.py::test_narrow_heads_through_the_fused_kernel[2-8-512-77-40]" tests/test_kernels_gpu.py -q -x > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"; tail -6 gpurun_out/sanitizer_memcheck.log
