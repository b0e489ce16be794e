# round-2 GPU call 32: full GPU suite after the host-side changes, bench
mkdir -p gpurun_out/r2
(timeout 1500 python -m pytest tests -q -m gpu --timeout 300 --tb=short 2>&1 | tail -6) > gpurun_out/r2/c32_tests.txt
(timeout 400 python bench.py --no-cpu-baseline 2>/dev/null) > gpurun_out/r2/c32_bench.json
cat gpurun_out/r2/c32_tests.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2/c32_bench.json'))
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'seq', {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['sequential'].items() if k!='note'})
PY
