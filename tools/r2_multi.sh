# usage: bash tools/r2_multi.sh N   (run under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
nvidia-smi -L | head -8
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/r2_multi_test_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_multi_test_n$N.log
  tail -5 gpurun_out/r2_multi_test_n$N.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench N=$N rc=$?"
tail -c 300 gpurun_out/r2_bench_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload train-cil --steps 10 --warmup 3 > gpurun_out/r2_bench_traincil_n$N.json 2> gpurun_out/r2_bench_traincil_n$N.err; echo "train-cil N=$N rc=$?"
tail -c 300 gpurun_out/r2_bench_traincil_n$N.err
