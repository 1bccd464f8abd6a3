mkdir -p gpurun_out
for k in 192 384 97; do
  export OAT_TC_DIRECT_K=$k
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_dk_$k.json 2> gpurun_out/r2_bench_dk_$k.err; echo "bench direct_k=$k rc=$?"
done
