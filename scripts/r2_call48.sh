# round-2 GPU call 48: 3 / 4 / 5 scans in flight after the GC freeze (end-to-end leg is host-limited: GPU busy 65 %)
mkdir -p gpurun_out/r2
for w in 3 4 5; do (timeout 400 python bench.py --no-cpu-baseline --inflight $w 2>/dev/null) > gpurun_out/r2/c48_bench_w$w.json
python - <<PY
import json
d=json.load(open('gpurun_out/r2/c48_bench_w$w.json'))
print('inflight $w: value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2))
PY
done
