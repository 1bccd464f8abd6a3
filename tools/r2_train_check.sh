mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q > gpurun_out/r2_train_coop_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_train_coop_test.log; tail -4 gpurun_out/r2_train_coop_test.log
python - <<'PY'
import torch, os, sys
sys.path.insert(0, os.getcwd())
# bitwise comparison of the cooperative decoder pass against the work-item functor
import subprocess
PY
python bench.py --workload train-dim --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_train_dim_coop.json 2> gpurun_out/r2_bench_train_dim_coop.err; echo "train-dim rc=$?"
OAT_TRAIN_COOP=0 python bench.py --workload train-dim --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_train_dim_nocoop.json 2> gpurun_out/r2_bench_train_dim_nocoop.err; echo "train-dim nocoop rc=$?"
