# round-2 GPU call 28: full-size c3 / c4 / c5 workloads through bench.py (coverage of BASELINE configs 3-5)
mkdir -p gpurun_out/r2
for w in c3 c4 c5; do (timeout 600 python bench.py --no-cpu-baseline --workload $w --steps 6 --warmup 3 2>gpurun_out/r2/c28_$w.err) > gpurun_out/r2/c28_$w.json; tail -2 gpurun_out/r2/c28_$w.err; done
python - <<'PY'
import json
for w in ('c3','c4','c5'):
    try:
        d=json.load(open('gpurun_out/r2/c28_%s.json'%w))
        print(w, d['metric'], 'value', round(d['value'],2), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'seq', round(d['sequential']['ms_per_step'],2), round(d['sequential']['e2e_ms_per_step'],2), 'proposals', d['config'].get('proposals'), d['config'].get('proposal_points'))
    except Exception as e:
        print(w, 'failed', repr(e)[:200])
PY
