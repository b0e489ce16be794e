# round-2 GPU call 30: bfs emit fix (edges into other components): ops tests incl. the new stress test, c5 debug run, c3/c5 bench lines
mkdir -p gpurun_out/r2
(timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -q -m gpu --timeout 300 --tb=short 2>&1 | tail -6) > gpurun_out/r2/c30_tests.txt
(timeout 300 python scripts/debug_workload.py c5 2>&1 | grep -v "^  bfs" | tail -12) > gpurun_out/r2/c30_c5.txt
for w in c3 c5; do (timeout 600 python bench.py --no-cpu-baseline --workload $w --steps 6 --warmup 3 2>gpurun_out/r2/c30_$w.err) > gpurun_out/r2/c30_$w.json; tail -2 gpurun_out/r2/c30_$w.err; done
cat gpurun_out/r2/c30_tests.txt gpurun_out/r2/c30_c5.txt
python - <<'PY'
import json
for w in ('c3','c5'):
    try:
        d=json.load(open('gpurun_out/r2/c30_%s.json'%w))
        print(w, d['metric'], 'value', round(d['value'],2), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'seq', round(d['sequential']['ms_per_step'],2), round(d['sequential']['e2e_ms_per_step'],2), 'proposals', d['config'].get('proposals'), d['config'].get('proposal_points'))
    except Exception as e:
        print(w, 'failed', repr(e)[:200])
PY
