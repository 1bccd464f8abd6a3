mkdir -p gpurun_out
python tools/tc_trace.py 4096 960 320 4 > gpurun_out/r2_trace2_f17proj.txt 2>&1
python tools/tc_trace.py 12544 384 64 4 > gpurun_out/r2_trace2_f8proj.txt 2>&1
python tools/tc_trace.py 12544 64 384 4 > gpurun_out/r2_trace2_f8exp.txt 2>&1
tail -9 gpurun_out/r2_trace2_f17proj.txt
timeout 900 python -m pytest tests/test_gpu_tc_gemm.py tests/test_gpu_parity.py -q > gpurun_out/r2_elect_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_elect_test.log
tail -4 gpurun_out/r2_elect_test.log
for w in 0 1; do
  export OAT_TC_WSPLIT=$w
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_elect_w$w.json 2> gpurun_out/r2_bench_elect_w$w.err; echo "bench wsplit=$w rc=$?"
done
