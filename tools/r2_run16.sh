mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc_gemm.py tests/test_gpu_fused.py -q > gpurun_out/r2_stackk_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_stackk_test.log; tail -3 gpurun_out/r2_stackk_test.log
OAT_TC_TS=1 timeout 600 python -m pytest tests/test_gpu_tc_gemm.py -q > gpurun_out/r2_stackk_ts_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_stackk_ts_test.log; tail -3 gpurun_out/r2_stackk_ts_test.log
for k in 192 99999 0 384; do
  export OAT_TC_STACK_K=$k
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_sk_$k.json 2> gpurun_out/r2_bench_sk_$k.err; echo "bench stack_k=$k rc=$?"
done
