#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --full-ft --batch 1 --steps 5 --warmup 3 --skip-cpu --skip-gpu-baseline > gpurun_out/bench_full_ft_graph.json 2> gpurun_out/bench_full_ft_graph.err
echo "bench full-ft (CUDA graph) exit $?"; cut -c1-200 gpurun_out/bench_full_ft_graph.json; tail -3 gpurun_out/bench_full_ft_graph.err; cp gpurun_out/gemm_by_shape.json gpurun_out/gemm_by_shape_full_ft.json
