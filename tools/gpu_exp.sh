#!/bin/bash
mkdir -p gpurun_out
{
OAT_EXPERIMENTAL=1 timeout 45 python -m pytest tests/test_gpu_experimental.py -x -q 2>&1 | tail -5
timeout 25 python tools/train_bench.py --no-cpu --graphs --steps 10 --warmup 3 2>&1 | tail -2
} > gpurun_out/exp.log 2>&1
tail -12 gpurun_out/exp.log | cut -c1-400
