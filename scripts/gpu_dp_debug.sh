#!/bin/bash
mkdir -p gpurun_out
for gmode in 1; do
  echo "=== DP_DEBUG_GRAPH=$gmode"
  DP_DEBUG_GRAPH=$gmode NCCL_DEBUG=WARN timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port $((29520 + gmode)) scripts/dp_debug.py > gpurun_out/dp_debug_$gmode.log 2>&1
  echo "exit $?"
  grep -a "rank \|Error\|error\|Traceback\|File \"/root\|NCCL WARN" gpurun_out/dp_debug_$gmode.log | head -40
done

echo "=== full bench, 2 GPUs"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "exit $?"; cut -c1-1200 gpurun_out/bench_2gpu.json; grep -a "Error\|error" gpurun_out/bench_2gpu.err | head -5
