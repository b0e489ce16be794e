# round-2 GPU call 38: evidence refresh after the ball-query / BFS changes: full GPU suite, smoke, bench (full line), launch list, ncu --set full of the grouping kernels
mkdir -p gpurun_out/r2
(timeout 1500 python -m pytest tests -q -m gpu --timeout 300 --tb=short 2>&1 | tail -4) > gpurun_out/r2/c38_tests.txt
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) > gpurun_out/r2/c38_smoke.txt
(timeout 600 python bench.py 2>gpurun_out/r2/c38_bench.err) > gpurun_out/r2/c38_bench.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2/c38_launches.csv python scripts/one_step.py 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k 'regex:bq_query|bfs_frontier|bfs_emit2' -c 16 -f -o gpurun_out/r2/c38_prof_grp python scripts/one_step.py 1 > /dev/null 2>&1
ncu -i gpurun_out/r2/c38_prof_grp.ncu-rep --page raw --csv > gpurun_out/r2/c38_prof_grp_raw.csv 2>/dev/null
rm -f gpurun_out/r2/c38_prof_grp.ncu-rep
cat gpurun_out/r2/c38_tests.txt gpurun_out/r2/c38_smoke.txt; python scripts/launch_summary.py gpurun_out/r2/c38_launches.csv 2>/dev/null | head -8
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2/c38_bench.json'))
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'seq', {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['sequential'].items() if k!='note'})
print(d['roofline']['frac'], d.get('cpu_baseline',{}).get('value'))
PY
