mkdir -p gpurun_out
for b in 0 80 64; do
  export OAT_TC_BN_SHALLOW=$b
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_shallow_$b.json 2> gpurun_out/r2_bench_shallow_$b.err; echo "bench shallow=$b rc=$?"
done
export OAT_TC_BN_SHALLOW=80
timeout 600 python -m pytest tests/test_gpu_tc_gemm.py tests/test_gpu_fused.py -q > gpurun_out/r2_shallow_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_shallow_test.log; tail -3 gpurun_out/r2_shallow_test.log
