mkdir -p gpurun_out
for k in 160 512; do
  export OAT_TC_BN_SHALLOW_K=$k
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_shk_$k.json 2> gpurun_out/r2_bench_shk_$k.err; echo "bench shallow_k=$k rc=$?"
done
