# round-2 GPU call 24: ncu evidence for round 2: launch list of one step, --set full of spconv_ss_kernel (levels 0-3) and the grouping kernels, DRAM bytes of every conv launch
mkdir -p gpurun_out/r2
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2/c24_launches.csv python scripts/one_step.py 1 > /dev/null 2>&1
timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,l1tex__t_bytes.sum --clock-control none --profile-from-start off -k regex:spconv_ --csv --log-file gpurun_out/r2/c24_conv_metrics.csv python scripts/one_step.py 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:spconv_ss_kernel -c 20 -f -o gpurun_out/r2/c24_prof_ss python scripts/one_step.py 1 > /dev/null 2>&1
ncu -i gpurun_out/r2/c24_prof_ss.ncu-rep --page raw --csv > gpurun_out/r2/c24_prof_ss_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k 'regex:bq_query|bfs_propagate|bfs_emit2|ge_' -c 12 -f -o gpurun_out/r2/c24_prof_grp python scripts/one_step.py 1 > /dev/null 2>&1
ncu -i gpurun_out/r2/c24_prof_grp.ncu-rep --page raw --csv > gpurun_out/r2/c24_prof_grp_raw.csv 2>/dev/null
ls -la gpurun_out/r2/ | grep c24
