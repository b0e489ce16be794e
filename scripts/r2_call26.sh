# round-2 GPU call 26 (2 GPUs): torchrun path of bench.py with scans in flight + core affinity
mkdir -p gpurun_out/r2
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 12 --warmup 3 2>gpurun_out/r2/c26_bench2.err) > gpurun_out/r2/c26_bench2.json
tail -5 gpurun_out/r2/c26_bench2.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2/c26_bench2.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'],'per_rank',d.get('per_rank'),'seq',d['sequential'])
PY
nproc; python -c "import os;print(len(os.sched_getaffinity(0)))"; numactl -H 2>/dev/null | head -5; nvidia-smi topo -m 2>/dev/null | head -12
