# round-2 GPU call 27: two MMA-issuing warps in spconv_ss_kernel: conv parity (both kernels), per-level A/B (product + counters), bench
mkdir -p gpurun_out/r2
(timeout 600 python -m pytest tests/test_gpu_spconv.py -q -m gpu --timeout 120 --tb=line 2>&1 | tail -6) > gpurun_out/r2/c27_tests.txt
(SS_LEVELS=0,1,2,3,4 timeout 200 python scripts/ss_timeline.py --no-tl 2>&1 | tail -8) > gpurun_out/r2/c27_ss_ab_product.txt
(SS_LEVELS=0,2,3 timeout 200 python scripts/ss_timeline.py 2>&1 | tail -12) > gpurun_out/r2/c27_ss_ab_tl.txt
(timeout 400 python bench.py --no-cpu-baseline 2>/dev/null) > gpurun_out/r2/c27_bench.json
cat gpurun_out/r2/c27_tests.txt gpurun_out/r2/c27_ss_ab_product.txt gpurun_out/r2/c27_ss_ab_tl.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2/c27_bench.json'))
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'seq', {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['sequential'].items() if k!='note'})
print({k:round(v,3) for k,v in d['stage_ms'].items() if 'conv' in k})
PY
