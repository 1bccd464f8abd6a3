set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest_1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputest_1.log
tail -5 gpurun_out/r2_gputest_1.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_a.json 2> gpurun_out/r2_bench_n1_a.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench_n1_a.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_a.json 2> gpurun_out/r2_bench_ref_a.err; echo "ref rc=$?"
python bench.py --workload train-dim --steps 10 --warmup 3 > gpurun_out/r2_bench_train_dim_a.json 2> gpurun_out/r2_bench_train_dim_a.err; echo "train-dim rc=$?"
python bench.py --workload train-cil --steps 10 --warmup 3 > gpurun_out/r2_bench_train_cil_a.json 2> gpurun_out/r2_bench_train_cil_a.err; echo "train-cil rc=$?"
tail -c 400 gpurun_out/r2_bench_train_cil_a.err
