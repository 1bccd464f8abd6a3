mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_gputest_3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputest_3.log
tail -12 gpurun_out/r2_gputest_3.log
timeout 300 python tools/tc_trace.py 12544 384 64 4 > gpurun_out/r2_trace6_f8proj.txt 2>&1
tail -8 gpurun_out/r2_trace6_f8proj.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; echo "bench rc=$?"
