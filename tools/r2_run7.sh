mkdir -p gpurun_out
python tools/tc_trace.py 4096 960 320 4 > gpurun_out/r2_trace_f17proj.txt 2>&1
python tools/tc_trace.py 4096 160 960 4 > gpurun_out/r2_trace_f15exp.txt 2>&1
python tools/tc_trace.py 12544 384 64 4 > gpurun_out/r2_trace_f8proj.txt 2>&1
OAT_TC_WSPLIT=0 OAT_TC_DIRECT=0 python tools/tc_trace.py 4096 960 320 4 > gpurun_out/r2_trace_f17proj_old.txt 2>&1
tail -12 gpurun_out/r2_trace_f17proj.txt
