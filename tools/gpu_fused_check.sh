#!/bin/bash
# GPU call: fused encoder front — probe, parity, sanitizer, timing per mask, ncu of the fused kernels.
mkdir -p gpurun_out
{
echo "=== probe"; timeout 100 python tools/sanitize_fused.py 2>&1 | tail -5
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "PROBE FAILED"; else
echo "=== pytest fused"; timeout 300 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -15
echo "=== memcheck"; timeout 200 compute-sanitizer --tool memcheck python tools/sanitize_fused.py 2>&1 | tail -12
echo "=== racecheck"; timeout 240 compute-sanitizer --tool racecheck python tools/sanitize_fused.py 2>&1 | tail -12
echo "=== fusion bench"; timeout 200 python tools/fusion_bench.py 2>&1 | tail -10
echo "=== ncu fused kernels"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"front_kernel|expand_dw" -c 4 -f -o gpurun_out/r1_fused python tools/fusion_bench.py --once 15 2>&1 | tail -3
ncu -i gpurun_out/r1_fused.ncu-rep --page raw --csv > gpurun_out/r1_fused_raw.csv 2>/dev/null
fi
} > gpurun_out/fused_check.log 2>&1
tail -70 gpurun_out/fused_check.log
