# round-2 GPU call 18: trace of the pair-stage ss kernel (levels 0, 2), whole-step A/B of the kernel choice (all tc / heuristic / all ss)
mkdir -p gpurun_out/r2
(SS_LEVELS=0,2 timeout 200 python scripts/ss_timeline.py --trace 2>&1 | tail -130) > gpurun_out/r2/c18_trace.txt
for m in 0 1; do (SGB_CONV_SS=$m timeout 300 python bench.py --no-cpu-baseline 2>/dev/null) > gpurun_out/r2/c18_bench_ss$m.json; done
python - <<'PY'
import json
for m in (0,1):
    d=json.load(open('gpurun_out/r2/c18_bench_ss%d.json'%m))
    print('SGB_CONV_SS=%d'%m, 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), {k:v for k,v in d['stage_ms'].items() if 'spconv' in k or 'pack' in k})
PY
cat gpurun_out/r2/c18_trace.txt
