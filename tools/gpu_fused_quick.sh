#!/bin/bash
# GPU call: quick parity + timing of the fused kernels (fusion bench, per-kernel ncu durations).
mkdir -p gpurun_out
{
echo "=== pytest fused"; timeout 300 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -4
echo "=== fusion bench"; timeout 200 python tools/fusion_bench.py 2>&1 | tail -8
echo "=== per-kernel durations (ncu, cold)"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"expand_dw|front_kernel" -c 4 --csv python tools/fusion_bench.py --once 15 2>&1 | grep -o 'ExpandDw[A-Za-z]*<[^>]*>\|front_kernel\|"gpu__time_duration.sum","[a-z]*","[0-9.,]*"' | paste - - | cut -c1-160
} > gpurun_out/fused_quick.log 2>&1
tail -40 gpurun_out/fused_quick.log
