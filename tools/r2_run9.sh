mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc_gemm.py tests/test_gpu_parity.py tests/test_gpu_fused.py -q > gpurun_out/r2_stack_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_stack_test.log
tail -6 gpurun_out/r2_stack_test.log
python tools/tc_trace.py 4096 960 320 4 > gpurun_out/r2_trace3_f17proj.txt 2>&1
python tools/tc_trace.py 12544 384 64 4 > gpurun_out/r2_trace3_f8proj.txt 2>&1
python tools/tc_trace.py 12544 64 384 4 > gpurun_out/r2_trace3_f8exp.txt 2>&1
tail -8 gpurun_out/r2_trace3_f17proj.txt
export OAT_TC_WSPLIT=0
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_stack.json 2> gpurun_out/r2_bench_stack.err; echo "bench rc=$?"
OAT_TC_DIRECT=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_stack_d0.json 2> gpurun_out/r2_bench_stack_d0.err; echo "bench rc=$?"
OAT_TC_DIRECT=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_stack_d1.json 2> gpurun_out/r2_bench_stack_d1.err; echo "bench rc=$?"
