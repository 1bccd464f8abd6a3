mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_tc_gemm.py -q -k "expand_dw_epilogue or tc_gemm or gemm" > gpurun_out/r2_direct_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_direct_test.log
tail -8 gpurun_out/r2_direct_test.log
for d in 0 1 auto; do
  if [ $d = auto ]; then unset OAT_TC_DIRECT; else export OAT_TC_DIRECT=$d; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_direct_$d.json 2> gpurun_out/r2_bench_direct_$d.err; echo "bench direct=$d rc=$?"
done
