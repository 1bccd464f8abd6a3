mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py -q -k "expand_dw_epilogue" > gpurun_out/r2_dwepi_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_dwepi_test.log
tail -30 gpurun_out/r2_dwepi_test.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_n1_c.json 2> gpurun_out/r2_bench_n1_c.err; echo "bench rc=$?"
OAT_FUSE=30 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_n1_c_fuse30.json 2> gpurun_out/r2_bench_n1_c_fuse30.err; echo "bench30 rc=$?"
