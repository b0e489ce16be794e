"""GPU occupancy with scans in flight: CUPTI activity records of 12 device-resident scans run 3 at a time -> fraction of the
wall time with at least one kernel running, time with >= 2 kernels running, per-kernel totals.
Usage: python scripts/inflight_trace.py [workers]"""
import sys
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, '.')
from softgroup_b200 import harness, ops, synth  # noqa: E402
from softgroup_b200.configs import model_cfg  # noqa: E402
from softgroup_b200.model import SoftGroup  # noqa: E402

workers = int(sys.argv[1]) if len(sys.argv) > 1 else 3
E2E = len(sys.argv) > 2 and sys.argv[2] == 'e2e'  # pinned host batch -> result dict instead of device-resident inputs
torch.set_num_threads(1)
torch.manual_seed(0)
model = SoftGroup(**model_cfg('scannet')).cuda().eval()
scan = synth.make_scan('c2_scannet', seed=0)
hb = harness.to_host_batch(scan)
inj = harness.pointwise_injection(scan, sigma=0.03, seed=0)
dev = harness.device_batch(hb)


def step(_):
    if E2E:
        harness.run_scan(model, hb, inject_pointwise=inj)
        return None
    vc, v2p, p2v = ops.voxelization_idx(dev['coords'], 1)
    d = {k: v for k, v in dev.items() if k not in ('coords', 'voxel_coords', 'v2p_map', 'p2v_map')}
    model.forward_test(device_only=True, inject_pointwise=inj, voxel_coords=vc, v2p_map=v2p, p2v_map=p2v, **d)
    return None


with torch.no_grad():
    for _ in range(3):
        step(0)
pipe = harness.ScanPipeline(model, workers=workers, freeze_gc=True)
pipe.map(step, range(2 * workers))
torch.cuda.synchronize()
N = 12
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    pipe.map(step, range(N))
    torch.cuda.synchronize()
evs = [(e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort()
t0, t1 = evs[0][0], max(e[1] for e in evs)
pts = []
for s, e, _ in evs:
    pts.append((s, 1))
    pts.append((e, -1))
pts.sort()
cur, last, busy1, busy2 = 0, t0, 0.0, 0.0
for t, d in pts:
    if cur >= 1:
        busy1 += t - last
    if cur >= 2:
        busy2 += t - last
    cur += d
    last = t
wall = t1 - t0
tot = defaultdict(float)
for s, e, n in evs:
    tot[n.split('(')[0][-60:]] += e - s
print('%d scans, %d in flight: wall %.2f ms per scan; GPU busy (>=1 activity) %.1f %%, >=2 concurrent %.1f %%; sum of activity durations %.2f ms per scan'
      % (N, workers, wall / N / 1e3, 100 * busy1 / wall, 100 * busy2 / wall, sum(tot.values()) / N / 1e3))
for n, v in sorted(tot.items(), key=lambda x: -x[1])[:12]:
    print('  %8.3f ms per scan  %s' % (v / N / 1e3, n))
