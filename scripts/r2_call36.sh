# round-2 GPU call 36: is the in-flight throughput host-bound? 3 / 4 / 6 scans in flight
mkdir -p gpurun_out/r2
for w in 3 4 6; do (timeout 400 python bench.py --no-cpu-baseline --inflight $w 2>/dev/null) > gpurun_out/r2/c36_bench_w$w.json
python - <<PY
import json
d=json.load(open('gpurun_out/r2/c36_bench_w$w.json'))
print('inflight $w: value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'seq', {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['sequential'].items() if k!='note'})
PY
done
