"""CPU, world_size 2 over gloo: the N>1 plumbing of bench.py (rank-local work, MAX-reduction of the timing vector,
rank 0 reports, other ranks of the reference arm exit without work). No GPU involved."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
# each rank times its own scans; the job time is the max over ranks, throughput = world * steps / max
t = torch.tensor([10.0 * (rank + 1), 20.0 + rank], dtype=torch.float64)
dist.barrier()
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert t.tolist() == [10.0 * world, 20.0 + world - 1], t
from softgroup_b200 import synth
a = synth.make_scan('c1_plumbing', seed=rank)
b = synth.make_scan('c1_plumbing', seed=rank)
assert (a['coords'] == b['coords']).all()
if rank == 0:
    print('OK value', world * 5 / (t[0].item() / 1e3))
dist.destroy_process_group()
''' % ROOT


def test_gloo_world2_timing_reduction(tmp_path):
    f = tmp_path / 'w.py'
    f.write_text(SCRIPT)
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                          '--master-addr', '127.0.0.1', '--master-port', '29571', str(f)], capture_output=True,
                         text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert 'OK value' in out.stdout


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2',
                          '--steps', '1', '--warmup', '0'], capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ''
