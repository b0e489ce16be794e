# round-2 GPU call 46: the committed bench line (by_kernel fractions), CPU-side quick checks on the box
mkdir -p gpurun_out/r2
(timeout 600 python bench.py 2>gpurun_out/r2/c46_bench.err) > gpurun_out/r2/c46_bench.json
tail -2 gpurun_out/r2/c46_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2/c46_bench.json'))
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), 'seq', round(d['sequential']['ms_per_step'],2), round(d['sequential']['e2e_ms_per_step'],2))
for k,v in d['roofline']['by_kernel'].items(): print('  %-30s %6.3f ms  %7.1f GB/s  frac %.3f %s' % (k, v['ms_per_step'], v['gbs'], v['frac_of_hbm_peak'], v.get('frac_of_issue_floor','')))
PY
