# round-2 GPU call 40: spread of the sequential steps (min / median / max) in two bench runs
mkdir -p gpurun_out/r2
for i in 1 2; do (timeout 400 python bench.py --no-cpu-baseline 2>/dev/null) > gpurun_out/r2/c40_bench_$i.json
python - <<PY
import json
d=json.load(open('gpurun_out/r2/c40_bench_$i.json'))
print('run $i: value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'seq', {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['sequential'].items() if k!='note'})
PY
done
