# round-2 GPU call 49 (4 GPUs): torchrun path at N=4 (NVML core sets: four ranks on one NUMA node)
mkdir -p gpurun_out/r2
(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 10 --warmup 3 2>gpurun_out/r2/c49_bench4.err) > gpurun_out/r2/c49_bench4.json
tail -3 gpurun_out/r2/c49_bench4.err
python - <<'PY'
import json
s=[l for l in open('gpurun_out/r2/c49_bench4.json') if l.startswith('{')][0]
d=json.loads(s)
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'n_gpus',d['n_gpus'],'e2e',round(d['e2e']['value'],1),round(d['e2e']['ms_per_step'],2),'\nper_rank',d.get('per_rank'))
PY
