# round-2 GPU call 11: tcgen05.mma issue-rate patterns + ncu --set full of the conv kernel (levels 0-4) and the grouping kernels
mkdir -p gpurun_out/r2
timeout 120 scripts/experiments/build/umma_rate2 > gpurun_out/r2/c11_umma_rate2.txt 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:spconv_tc_kernel -c 26 -f -o gpurun_out/r2/c11_prof_tc python scripts/one_step.py 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k 'regex:bq_query|bfs_propagate|bfs_emit2|rb_subm3|rb_down|vox_' -c 40 -f -o gpurun_out/r2/c11_prof_ops python scripts/one_step.py 1 > /dev/null 2>&1
cat gpurun_out/r2/c11_umma_rate2.txt | head -150; ls -la gpurun_out/r2
