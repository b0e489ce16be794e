# round-2 GPU call 44: warp-aggregated member cursors, float4 row gather in clusters_voxelization: model/ops/golden tests, launch list, bench
mkdir -p gpurun_out/r2
(timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_forward_golden.py tests/test_gpu_reference_model.py -q -m gpu --timeout 300 --tb=short 2>&1 | tail -4) > gpurun_out/r2/c44_tests.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2/c44_launches.csv python scripts/one_step.py 1 > /dev/null 2>&1
(timeout 400 python bench.py --no-cpu-baseline 2>/dev/null) > gpurun_out/r2/c44_bench.json
cat gpurun_out/r2/c44_tests.txt; python scripts/launch_summary.py gpurun_out/r2/c44_launches.csv 2>/dev/null | head -22
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2/c44_bench.json'))
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'seq', {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['sequential'].items() if k!='note'})
PY
