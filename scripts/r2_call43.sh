# round-2 GPU call 43: ball query with per-cell id ranges (one global pass and two barriers less per work item): tests, kernel times, bench
mkdir -p gpurun_out/r2
(timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_vs_reference.py tests/test_gpu_model.py -q -m gpu --timeout 300 --tb=short 2>&1 | tail -4) > gpurun_out/r2/c43_tests.txt
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --profile-from-start off -k regex:bq_ --csv --log-file gpurun_out/r2/c43_bq.csv python scripts/one_step.py 1 > /dev/null 2>&1
(timeout 400 python bench.py --no-cpu-baseline 2>/dev/null) > gpurun_out/r2/c43_bench.json
cat gpurun_out/r2/c43_tests.txt; grep -v "^==" gpurun_out/r2/c43_bq.csv | cut -d, -f5,13- | tail -14
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2/c43_bench.json'))
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'seq', {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['sequential'].items() if k!='note'})
print('  ', {k:round(v,3) for k,v in d['stage_ms'].items() if 'ball' in k})
PY
