mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_tc_gemm.py tests/test_gpu_parity.py -q > gpurun_out/r2_wsplit_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_wsplit_test.log
tail -6 gpurun_out/r2_wsplit_test.log
for w in 0 1; do
  export OAT_TC_WSPLIT=$w
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_wsplit_$w.json 2> gpurun_out/r2_bench_wsplit_$w.err; echo "bench wsplit=$w rc=$?"
done
export OAT_TC_WSPLIT=1 OAT_TC_DIRECT=0
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_wsplit_1_direct0.json 2> gpurun_out/r2_bench_wsplit_1_direct0.err; echo "bench wsplit=1 direct=0 rc=$?"
