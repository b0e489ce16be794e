"""Distribution of the end-to-end time of single sequential calls and of scans in flight (host allocator / thread effects).
Usage: python scripts/e2e_jitter.py"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from softgroup_b200 import harness, synth  # noqa: E402
from softgroup_b200.configs import model_cfg  # noqa: E402
from softgroup_b200.model import SoftGroup  # noqa: E402

torch.manual_seed(0)
model = SoftGroup(**model_cfg('scannet')).cuda().eval()
scan = synth.make_scan('c2_scannet', seed=0)
hb = harness.to_host_batch(scan)
inj = harness.pointwise_injection(scan, sigma=0.03, seed=0)
with torch.no_grad():
    for _ in range(3):
        harness.run_scan(model, hb, inject_pointwise=inj)
    torch.cuda.synchronize()
    ts = []
    for _ in range(30):
        t0 = time.perf_counter()
        r = harness.run_scan(model, hb, inject_pointwise=inj)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print('sequential e2e ms: median %.2f min %.2f max %.2f' % (np.median(ts), min(ts), max(ts)), ['%.1f' % t for t in ts])
for w in (2, 3, 4):
    pipe = harness.ScanPipeline(model, workers=w)
    pipe.map(lambda _: (harness.run_scan(model, hb, inject_pointwise=inj), None)[1], range(2 * w))
    torch.cuda.synchronize()
    for rep in range(3):
        t0 = time.perf_counter()
        pipe.map(lambda _: (harness.run_scan(model, hb, inject_pointwise=inj), None)[1], range(24))
        torch.cuda.synchronize()
        print('in flight %d: %.2f ms per scan' % (w, (time.perf_counter() - t0) * 1e3 / 24))
    pipe.close()
