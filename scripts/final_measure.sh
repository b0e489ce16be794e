mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/r1f_tests.txt
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r1f_bench.json 2> gpurun_out/r1f_bench.err
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r1f_launches.csv python scripts/one_step.py 1 > /dev/null 2>&1
timeout 150 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:spconv_tc --csv --log-file gpurun_out/r1f_dram_tc.csv python scripts/one_step.py 1 > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:spconv_tc_kernel -c 24 -f -o gpurun_out/r1f_prof_tc python scripts/one_step.py 1 > /dev/null 2>&1
timeout 150 ncu --set full --clock-control none --import-source on --profile-from-start off -k 'regex:bq_query|bfs_propagate|bfs_emit2|rb_subm3|act_split' -c 8 -f -o gpurun_out/r1f_prof_ops python scripts/one_step.py 1 > /dev/null 2>&1
cat gpurun_out/r1f_tests.txt; head -c 600 gpurun_out/r1f_bench.json; echo; tail -3 gpurun_out/r1f_bench.err; ls -la gpurun_out | tail -8
