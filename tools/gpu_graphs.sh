#!/bin/bash
# GPU call: CUDA-graph replay of the encoder stage — tests + bench with and without graphs.
mkdir -p gpurun_out
{
echo "=== pytest graphs"; timeout 300 python -m pytest tests/test_gpu_graphs.py -x -q 2>&1 | tail -12
echo "=== bench graphs"; timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_graphs.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['stages_ms'], d['gpu_launches'], d['e2e']['value'], d['e2e']['blocking_call_ms'])"
echo "=== bench no graphs"; timeout 300 python bench.py --no-cpu-baseline --no-cuda-graphs 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['stages_ms'], d['gpu_launches'], d['e2e']['value'], d['e2e']['blocking_call_ms'])"
} > gpurun_out/graphs.log 2>&1
tail -30 gpurun_out/graphs.log
