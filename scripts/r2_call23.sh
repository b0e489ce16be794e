# round-2 GPU call 23: e2e jitter (sequential / in flight), bench with 2 and 3 scans in flight
mkdir -p gpurun_out/r2
(timeout 300 python scripts/e2e_jitter.py 2>&1 | tail -14) > gpurun_out/r2/c23_jitter.txt
for w in 2 3; do (timeout 400 python bench.py --no-cpu-baseline --inflight $w 2>gpurun_out/r2/c23_bench_w$w.err) > gpurun_out/r2/c23_bench_w$w.json; done
cat gpurun_out/r2/c23_jitter.txt
python - <<'PY'
import json
for f in ('w2','w3'):
    try:
        d=json.load(open('gpurun_out/r2/c23_bench_%s.json'%f))
        print(f, 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'seq', {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['sequential'].items() if k!='note'})
    except Exception as e:
        print(f, 'failed', e)
PY
