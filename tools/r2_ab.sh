# usage: bash tools/r2_ab.sh "<ENV=.. ENV=..>" name   -> gpurun_out/r2_ab_<name>.json (N=1 bench, no CPU arms)
mkdir -p gpurun_out
env $1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_ab_$2.json 2> gpurun_out/r2_ab_$2.err; echo "bench $2 rc=$?"
