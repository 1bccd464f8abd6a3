#!/bin/bash
# GPU call (2 GPUs): multi-rank tests + the N=2 bench line.
mkdir -p gpurun_out
{
echo "=== pytest multi"; timeout 400 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5
echo "=== bench N=2"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_n2.json
} > gpurun_out/n2.log 2>&1
tail -12 gpurun_out/n2.log
