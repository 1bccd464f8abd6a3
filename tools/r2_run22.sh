mkdir -p gpurun_out
python tools/tc_trace.py 4096 960 160 4 > gpurun_out/r2_trace7_f15proj.txt 2>&1
python tools/tc_trace.py 12544 96 576 4 > gpurun_out/r2_trace7_f12exp.txt 2>&1
python tools/tc_trace.py 12544 576 96 4 > gpurun_out/r2_trace7_f12proj.txt 2>&1
python tools/tc_trace.py 4096 320 1280 4 > gpurun_out/r2_trace7_last.txt 2>&1
for f in f15proj f12exp f12proj last; do echo == $f; head -1 gpurun_out/r2_trace7_$f.txt; tail -13 gpurun_out/r2_trace7_$f.txt; done
