"""Scratch probe (GPU box): calibrated synthetic checkpoint -> full forward, per-stage ms and grouping stats."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from softgroup_b200 import harness, synth  # noqa: E402
from softgroup_b200.configs import model_cfg  # noqa: E402
from softgroup_b200.model import SoftGroup  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else 'c2_scannet'
cfgname = {'c2_scannet': 'scannet', 'c3_s3dis': 's3dis', 'c4_kitti': 'kitti'}[shape]
torch.manual_seed(0)
model = SoftGroup(**model_cfg(cfgname)).cuda().eval()
scan = synth.make_scan(shape, seed=0)
hb = harness.to_host_batch(scan)
t = time.time()
inj = harness.pointwise_injection(scan, sigma=float(sys.argv[2]) if len(sys.argv) > 2 else 0.03)
model.profile_stages = True
with torch.no_grad():
    for it in range(3):
        torch.cuda.synchronize()
        t = time.time()
        ret = harness.run_scan(model, hb, inject_pointwise=inj)
        torch.cuda.synchronize()
        dt = (time.time() - t) * 1e3
        print('iter %d: %.1f ms  stages %s' % (it, dt, {k: round(v, 2) for k, v in model.stage_ms.items()}))
    ret = harness.run_scan(model, hb, device_only=True, inject_pointwise=inj)
    po = ret['proposals_offset'].cpu().numpy()
    print('nProposal', len(po) - 1, 'sumNPoint', po[-1] if len(po) else 0, 'gt instances', len(scan['instance_pointnum']))
    if len(po) > 1:
        sz = np.diff(po)
        print('proposal sizes: min %d med %d max %d' % (sz.min(), np.median(sz), sz.max()))
    print('nActive stats: see profiler'); print('pred_instances', len(harness.run_scan(model, hb, inject_pointwise=inj)['pred_instances']))
