# Final single-GPU evidence of round 2 (run under gpurun, 1 GPU).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_final_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_final_gputest.log
tail -6 gpurun_out/r2_final_gputest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err; echo "ref rc=$?"
python bench.py --workload train-dim --steps 10 --warmup 3 > gpurun_out/r2_final_bench_train_dim.json 2> gpurun_out/r2_final_bench_train_dim.err; echo "train-dim rc=$?"
python bench.py --workload train-cil --steps 10 --warmup 3 > gpurun_out/r2_final_bench_train_cil.json 2> gpurun_out/r2_final_bench_train_cil.err; echo "train-cil rc=$?"
# every launch of one step: device time + DRAM traffic + tensor-pipe activity (second of two steps)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none -k regex:"oat" -s 59 -c 59 --csv --log-file gpurun_out/r2_step_launches.csv python tools/r2_step_only.py 2 > gpurun_out/r2_step_launches.log 2>&1
tail -1 gpurun_out/r2_step_launches.log
