#!/bin/bash
# GPU call (4 GPUs): multi-rank tests + the N=2 and N=4 bench lines.
mkdir -p gpurun_out
{
echo "=== pytest multi"; timeout 400 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5
for n in 2 4; do
echo "=== bench N=$n"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_n$n.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['stages_ms'], d['e2e']['value'])"
done
} > gpurun_out/n4.log 2>&1
tail -12 gpurun_out/n4.log
