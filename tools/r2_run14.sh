mkdir -p gpurun_out
export OAT_TC_BN_SHALLOW=64
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none -k regex:"tc_pw_gemm|dw|stem|expand|pool|merger|front" -s 55 -c 55 --csv --log-file gpurun_out/r2_encoder_launches.csv python tools/r2_encode_only.py 2 > gpurun_out/r2_encoder_launches.log 2>&1
tail -2 gpurun_out/r2_encoder_launches.log
