#!/bin/bash
# GPU call: training step with the tiled pointwise GEMMs — gradient parity + step time A/B.
mkdir -p gpurun_out
{
echo "=== pytest train (tiled)"; OAT_TRAIN_TILED=1 timeout 120 python -m pytest tests/test_gpu_train.py -x -q 2>&1 | tail -4
echo "=== train bench (tiled)"; OAT_TRAIN_TILED=1 timeout 60 python tools/train_bench.py --no-cpu 2>&1 | tail -2
echo "=== train bench (functors)"; OAT_TRAIN_TILED=0 timeout 60 python tools/train_bench.py --no-cpu 2>&1 | tail -2
} > gpurun_out/train_tiled.log 2>&1
tail -20 gpurun_out/train_tiled.log | cut -c1-500
