mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q > gpurun_out/r2_train_coop_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_train_coop_test.log; tail -4 gpurun_out/r2_train_coop_test.log
python bench.py --workload train-dim --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_train_dim_coop.json 2> gpurun_out/r2_bench_train_dim_coop.err; echo "train-dim rc=$?"
python bench.py --workload train-cil --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_train_cil_coop.json 2> gpurun_out/r2_bench_train_cil_coop.err; echo "train-cil rc=$?"
