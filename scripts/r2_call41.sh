# round-2 GPU call 41 (2 GPUs): torchrun path of the final bench.py (NVML core sets, scans in flight, GC freeze)
mkdir -p gpurun_out/r2
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 12 --warmup 3 2>gpurun_out/r2/c41_bench2.err) > gpurun_out/r2/c41_bench2.json
tail -3 gpurun_out/r2/c41_bench2.err
python - <<'PY'
import json
s=[l for l in open('gpurun_out/r2/c41_bench2.json') if l.startswith('{')][0]
d=json.loads(s)
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'n_gpus',d['n_gpus'],'e2e',d['e2e'],'\nper_rank',d.get('per_rank'),'\nseq',d['sequential'])
PY
python - <<'PY'
import os
print('affinity of this shell', len(os.sched_getaffinity(0)))
PY
