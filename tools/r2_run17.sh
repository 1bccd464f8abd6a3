mkdir -p gpurun_out
for b in 0 128 96; do
  export OAT_TC_BN_SHALLOW=$b
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_sh2_$b.json 2> gpurun_out/r2_bench_sh2_$b.err; echo "bench shallow=$b rc=$?"
done
