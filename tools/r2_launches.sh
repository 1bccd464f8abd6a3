mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none -k regex:"aggregate_kernel|dw2_kernel|dw_kernel|dw_project_kernel|expand_dw|flow_tc2|merger_kernel|pool_kernel|stem_kernel|tc_pw_gemm|transform_visual" -c 400 --csv --log-file gpurun_out/r2_step_launches.csv python tools/r2_step_only.py 2 > gpurun_out/r2_step_launches.log 2>&1
tail -1 gpurun_out/r2_step_launches.log
# full-set capture of the dominant family's heaviest launches (features.18 and the 960->320 project) for profiles/
true
true
# memory checker over one small end-to-end step with the final kernels (default path, and the opt-in GEMM forms)
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py > gpurun_out/r2_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2_memcheck.log
OAT_FUSE=62 OAT_TC_TS=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py > gpurun_out/r2_memcheck_optin.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2_memcheck_optin.log
tail -3 gpurun_out/r2_memcheck.log; tail -3 gpurun_out/r2_memcheck_optin.log
