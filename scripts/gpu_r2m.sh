#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --full-ft --batch 1 --steps 5 --warmup 3 --skip-cpu --skip-gpu-baseline > gpurun_out/bench_full_ft_graph.json 2> gpurun_out/bench_full_ft_graph.err
echo "bench full-ft (CUDA graph) exit $?"; cut -c1-200 gpurun_out/bench_full_ft_graph.json; tail -3 gpurun_out/bench_full_ft_graph.err; cp gpurun_out/gemm_by_shape.json gpurun_out/gemm_by_shape_full_ft.json
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-gpu-baseline --skip-roofline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default exit $?"; cut -c1-200 gpurun_out/bench_default.json
