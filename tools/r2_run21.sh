mkdir -p gpurun_out
for m in 0 4 7 13; do
  export OAT_DW_PX_MAX=$m
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_dwpx_$m.json 2> gpurun_out/r2_bench_dwpx_$m.err; echo "bench dwpx=$m rc=$?"
done
export OAT_DW_PX_MAX=13
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "goldens" > gpurun_out/r2_dwpx_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_dwpx_test.log; tail -3 gpurun_out/r2_dwpx_test.log
