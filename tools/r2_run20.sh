mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flow_tc2 -s 2 -c 2 -f -o gpurun_out/r2_flow python tools/r2_flow_only.py 2 > gpurun_out/r2_ncu_flow.log 2>&1
tail -2 gpurun_out/r2_ncu_flow.log; ls -la gpurun_out/r2_flow.ncu-rep
