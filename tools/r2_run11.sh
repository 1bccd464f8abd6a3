mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc_gemm.py -q -x > gpurun_out/r2_split_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_split_test.log
tail -3 gpurun_out/r2_split_test.log
timeout 300 python tools/tc_trace.py 4096 960 320 4 > gpurun_out/r2_trace5_f17proj.txt 2>&1
timeout 300 python tools/tc_trace.py 12544 64 384 4 > gpurun_out/r2_trace5_f8exp.txt 2>&1
timeout 300 python tools/tc_trace.py 12544 384 64 4 > gpurun_out/r2_trace5_f8proj.txt 2>&1
OAT_TC_TS=1 timeout 300 python tools/tc_trace.py 4096 960 320 4 > gpurun_out/r2_trace5_f17proj_ts.txt 2>&1
tail -8 gpurun_out/r2_trace5_f17proj.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_split.json 2> gpurun_out/r2_bench_split.err; echo "bench rc=$?"
OAT_TC_TS=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_split_ts.json 2> gpurun_out/r2_bench_split_ts.err; echo "bench rc=$?"
OAT_TC_DIRECT=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_split_d0.json 2> gpurun_out/r2_bench_split_d0.err; echo "bench rc=$?"
