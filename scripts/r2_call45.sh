# round-2 GPU call 45: final validation: full GPU suite, smoke(), full bench line (cpu baseline + reference GPU ops), reference arm
mkdir -p gpurun_out/r2
(timeout 1500 python -m pytest tests -q -m gpu --timeout 300 --tb=short 2>&1 | tail -3) > gpurun_out/r2/c45_tests.txt
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) > gpurun_out/r2/c45_smoke.txt
(timeout 600 python bench.py 2>gpurun_out/r2/c45_bench.err) > gpurun_out/r2/c45_bench.json
(timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null) > gpurun_out/r2/c45_bench_ref.json
cat gpurun_out/r2/c45_tests.txt gpurun_out/r2/c45_smoke.txt; tail -2 gpurun_out/r2/c45_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2/c45_bench.json'))
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', d['e2e'], '\nseq', d['sequential'], '\nlaunches', d['gpu_launches_per_step'], d['clocks'])
print({k:v for k,v in d['roofline'].items() if k not in ('by_kernel','traffic_source')})
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
r=json.load(open('gpurun_out/r2/c45_bench_ref.json')); print('ref arm', r['value'], r['impl'])
PY
