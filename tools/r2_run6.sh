mkdir -p gpurun_out
# one encode = 31 tc_pw_gemm launches; skip the first encode + 27 -> f17 expand, f17 project, features.18
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_pw_gemm -s 58 -c 3 -f -o gpurun_out/r2_gemm_late python tools/r2_encode_only.py 2 > gpurun_out/r2_ncu_gemm_late.log 2>&1
tail -3 gpurun_out/r2_ncu_gemm_late.log
ls -la gpurun_out/*.ncu-rep
