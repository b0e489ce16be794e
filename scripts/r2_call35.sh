# round-2 GPU call 35: bfs labelling as select + balanced push kernels: ops/model tests, launch list of one step, bench
mkdir -p gpurun_out/r2
(timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_reference_model.py -q -m gpu --timeout 300 --tb=short 2>&1 | tail -5) > gpurun_out/r2/c35_tests.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2/c35_launches.csv python scripts/one_step.py 1 > /dev/null 2>&1
(timeout 400 python bench.py --no-cpu-baseline 2>/dev/null) > gpurun_out/r2/c35_bench.json
cat gpurun_out/r2/c35_tests.txt
python scripts/launch_summary.py gpurun_out/r2/c35_launches.csv 2>/dev/null | head -14
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2/c35_bench.json'))
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'seq', {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['sequential'].items() if k!='note'})
print('  ', {k:round(v,3) for k,v in d['stage_ms'].items() if 'ball' in k or 'bfs' in k or 'pack' in k})
PY
