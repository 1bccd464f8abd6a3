#!/bin/bash
# GPU call: end-of-session validation — full GPU test suite, both bench arms, launch list.
mkdir -p gpurun_out
{
echo "=== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "=== smoke"; timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
echo "=== bench"; timeout 400 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_b200.json
echo "=== bench reference arm"; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference.json
echo "=== memcheck (end-to-end small step, default kernels)"; timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_small.py tcgen05x2 2>&1 | tail -4
echo "=== launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/launches.csv 2>&1 | head -20
} > gpurun_out/final.log 2>&1
tail -50 gpurun_out/final.log
