mkdir -p gpurun_out
for g in 0 2 1; do
  export OAT_ENC_GROUP=$g
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_grp_$g.json 2> gpurun_out/r2_bench_grp_$g.err; echo "bench group=$g rc=$?"
done
export OAT_ENC_GROUP=2 OAT_ENC_GROUP_SERIAL=1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_grp_2s.json 2> gpurun_out/r2_bench_grp_2s.err; echo "bench group=2 serial rc=$?"
unset OAT_ENC_GROUP_SERIAL
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused.py tests/test_gpu_graphs.py -q -x > gpurun_out/r2_grp_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_grp_test.log; tail -3 gpurun_out/r2_grp_test.log
unset OAT_ENC_GROUP
OAT_TC_BN_SHALLOW_K=512 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_shk_512.json 2> gpurun_out/r2_bench_shk_512.err; echo "bench shallow_k=512 rc=$?"
