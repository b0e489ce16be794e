# round-2 GPU call 50: instance dicts built in one comprehension: model / reference-model / golden tests, bench
mkdir -p gpurun_out/r2
(timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_reference_model.py tests/test_gpu_forward_golden.py -q -m gpu --timeout 300 --tb=short 2>&1 | tail -3) > gpurun_out/r2/c50_tests.txt
(timeout 400 python bench.py --no-cpu-baseline 2>/dev/null) > gpurun_out/r2/c50_bench.json
cat gpurun_out/r2/c50_tests.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2/c50_bench.json'))
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), 'seq', round(d['sequential']['ms_per_step'],2), round(d['sequential']['e2e_ms_per_step'],2))
PY
