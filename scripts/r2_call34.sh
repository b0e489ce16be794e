# round-2 GPU call 34: ball query inner loops on raw shared-memory addresses (count loop two blocks per trip): ops + reference-kernel parity tests, bench
mkdir -p gpurun_out/r2
(timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_vs_reference.py tests/test_gpu_model.py -q -m gpu --timeout 300 --tb=short 2>&1 | tail -6) > gpurun_out/r2/c34_tests.txt
(timeout 400 python bench.py --no-cpu-baseline 2>/dev/null) > gpurun_out/r2/c34_bench.json
(timeout 400 python bench.py --no-cpu-baseline --workload c2frag 2>/dev/null) > gpurun_out/r2/c34_bench_frag.json
cat gpurun_out/r2/c34_tests.txt
python - <<'PY'
import json
for f in ('c34_bench','c34_bench_frag'):
    d=json.load(open('gpurun_out/r2/%s.json'%f))
    print(f,'value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'seq', {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['sequential'].items() if k!='note'})
    print('  ', {k:round(v,3) for k,v in d['stage_ms'].items() if 'ball' in k or 'bfs' in k})
PY
