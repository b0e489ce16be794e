# round-2 GPU call 22: scans in flight (harness.ScanPipeline): equality test, bench with 1/2/3 in flight
mkdir -p gpurun_out/r2
(timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 300 --tb=short 2>&1 | tail -15) > gpurun_out/r2/c22_tests.txt
for w in 2 3; do (timeout 400 python bench.py --no-cpu-baseline --inflight $w 2>gpurun_out/r2/c22_bench_w$w.err) > gpurun_out/r2/c22_bench_w$w.json; done
(timeout 400 python bench.py --no-cpu-baseline --workload c2frag 2>/dev/null) > gpurun_out/r2/c22_bench_frag.json
cat gpurun_out/r2/c22_tests.txt
python - <<'PY'
import json
for f in ('w2','w3','frag'):
    try:
        d=json.load(open('gpurun_out/r2/c22_bench_%s.json'%f))
        print(f, 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'seq', {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['sequential'].items() if k!='note'}, d['clocks'])
    except Exception as e:
        print(f, 'failed', e)
PY
tail -5 gpurun_out/r2/c22_bench_w2.err
