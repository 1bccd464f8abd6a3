mkdir -p gpurun_out
export OAT_TC_TS=1
timeout 600 python -m pytest tests/test_gpu_tc_gemm.py -q -x > gpurun_out/r2_ts_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_ts_test.log
tail -15 gpurun_out/r2_ts_test.log
timeout 300 python tools/tc_trace.py 4096 960 320 4 > gpurun_out/r2_trace4_f17proj.txt 2>&1
timeout 300 python tools/tc_trace.py 12544 64 384 4 > gpurun_out/r2_trace4_f8exp.txt 2>&1
tail -8 gpurun_out/r2_trace4_f17proj.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_ts.json 2> gpurun_out/r2_bench_ts.err; echo "bench rc=$?"
