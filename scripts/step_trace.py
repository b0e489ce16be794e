"""Timeline of ONE device step from the CUPTI activity records (torch.profiler): every kernel/memcpy with start and
duration, and the idle gaps between consecutive GPU activities (where host syncs / launch latency show up).
Usage: python scripts/step_trace.py [workload-shape] > gpurun_out/step_trace.txt"""
import sys

import torch
from torch.profiler import ProfilerActivity, profile

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
flush = torch.empty(512 << 20, dtype=torch.uint8, device='cuda')


def step():
    vc, v2p, p2v = ops.voxelization_idx(dev['coords'], 1)
    d = {k: v for k, v in dev.items() if k not in ('coords', 'voxel_coords', 'v2p_map', 'p2v_map')}
    return model.forward_test(device_only=True, inject_pointwise=inj, voxel_coords=vc, v2p_map=v2p, p2v_map=p2v, **d)


with torch.no_grad():
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    flush.zero_()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()

evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
end_prev = t0
busy = 0.0
gaps = []
print('# start_us  dur_us  gap_before_us  name')
for e in evs:
    s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
    gap = e.time_range.start - end_prev
    print('%9.1f %8.1f %8.1f  %s' % (s, d, gap, e.name[:100]))
    if gap > 5:
        gaps.append((gap, s, e.name[:80]))
    busy += d
    end_prev = max(end_prev, e.time_range.end)
total = end_prev - t0
print('# total %.1f us, busy %.1f us, idle %.1f us over %d activities' % (total, busy, total - busy, len(evs)))
gaps.sort(reverse=True)
print('# largest gaps (us, at, before kernel):')
for g in gaps[:40]:
    print('#  %8.1f at %9.1f  %s' % g)
