# round-2 GPU call 33: bench with 1 / 8 / 128 intra-op threads (host-side noise of the in-flight legs)
mkdir -p gpurun_out/r2
for t in 1 8 128 1; do (SGB_BENCH_THREADS=$t timeout 400 python bench.py --no-cpu-baseline 2>/dev/null) > gpurun_out/r2/c33_bench_t$t.json
python - <<PY
import json
d=json.load(open('gpurun_out/r2/c33_bench_t$t.json'))
print('threads $t: value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'seq', {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['sequential'].items() if k!='note'})
PY
done
