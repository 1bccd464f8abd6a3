mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_gputest_2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputest_2.log
tail -15 gpurun_out/r2_gputest_2.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err; echo "bench rc=$?"
