# Final single-GPU validation of round 2 (run under gpurun, 1 GPU): what the driver runs at round end.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_final_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_final_gputest.log
tail -6 gpurun_out/r2_final_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_final_smoke.log; tail -2 gpurun_out/r2_final_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err; echo "ref rc=$?"
