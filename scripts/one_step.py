"""One warm-up + N device steps of the hot path (for ncu launch lists / captures). Usage: one_step.py [steps]"""
import sys

import torch

sys.path.insert(0, '.')
from softgroup_b200 import harness, ops, synth  # noqa: E402
from softgroup_b200.configs import model_cfg  # noqa: E402
from softgroup_b200.model import SoftGroup  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
torch.manual_seed(0)
model = SoftGroup(**model_cfg('scannet')).cuda().eval()
scan = synth.make_scan('c2_scannet', seed=0)
hb = harness.to_host_batch(scan)
inj = harness.pointwise_injection(scan, sigma=0.03, seed=0)
dev = harness.device_batch(hb)
with torch.no_grad():
    for it in range(1 + steps):
        if it == 1:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        vc, v2p, p2v = ops.voxelization_idx(dev['coords'], 1)
        d = {k: v for k, v in dev.items() if k not in ('coords', 'voxel_coords', 'v2p_map', 'p2v_map')}
        out = model.forward_test(device_only=True, inject_pointwise=inj, voxel_coords=vc, v2p_map=v2p, p2v_map=p2v, **d)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print('done', out['proposals_offset'].numel() - 1)
