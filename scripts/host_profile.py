"""cProfile of the host side of ONE device step (the Python/ctypes/torch work between kernel launches).
Usage: python scripts/host_profile.py > gpurun_out/host_profile.txt"""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, '.')
from softgroup_b200 import harness, ops, synth  # noqa: E402
from softgroup_b200.configs import model_cfg  # noqa: E402
from softgroup_b200.model import SoftGroup  # noqa: E402

torch.manual_seed(0)
model = SoftGroup(**model_cfg('scannet')).cuda().eval()
scan = synth.make_scan('c2_scannet', seed=0)
hb = harness.to_host_batch(scan)
inj = harness.pointwise_injection(scan, sigma=0.03, seed=0)
dev = harness.device_batch(hb)


def step():
    vc, v2p, p2v = ops.voxelization_idx(dev['coords'], 1)
    d = {k: v for k, v in dev.items() if k not in ('coords', 'voxel_coords', 'v2p_map', 'p2v_map')}
    return model.forward_test(device_only=True, inject_pointwise=inj, voxel_coords=vc, v2p_map=v2p, p2v_map=p2v, **d)


with torch.no_grad():
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        step()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        ts.append(((t1 - t0) * 1e3, (t2 - t0) * 1e3))
    print('host time of a step (ms) / until the GPU is done (ms):', ['%.2f / %.2f' % t for t in ts])
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    pr.disable()
st = pstats.Stats(pr, stream=sys.stdout)
st.sort_stats('cumulative').print_stats(45)
st.sort_stats('tottime').print_stats(30)
