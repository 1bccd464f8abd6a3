#!/bin/bash
# GPU call: fused encoder front — parity, sanitizer, timing per mask, ncu, full bench.
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
{
echo "=== pytest fused"; timeout 400 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -15
echo "=== memcheck"; timeout 200 compute-sanitizer --tool memcheck python tools/sanitize_fused.py 2>&1 | tail -12
echo "=== racecheck"; timeout 240 compute-sanitizer --tool racecheck python tools/sanitize_fused.py 2>&1 | tail -12
echo "=== fusion bench"; timeout 200 python tools/fusion_bench.py 2>&1 | tail -10
for s in 1 4; do OAT_FUSED_SPLITS=$s timeout 200 python tools/fusion_bench.py 2>&1 | grep "mask 15\|mask  [1248]" ; done
echo "=== bench fused"; OAT_FUSE=15 timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -2
echo "=== ncu fused kernels"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"front_kernel|expand_dw_kernel" -c 4 -f -o gpurun_out/r1_fused python tools/fusion_bench.py --once 15 2>&1 | tail -3
ncu -i gpurun_out/r1_fused.ncu-rep --page raw --csv > gpurun_out/r1_fused_raw.csv 2>/dev/null
} > gpurun_out/call1.log 2>&1
tail -60 gpurun_out/call1.log
