"""A/B of the off-by-default candidates (cooperative gather, wave-aware split-K, frontier BFS labelling) against the
validated defaults on the real device step (L2 flushed between steps); proposals must stay bit-identical."""
import ctypes
import sys

import torch

sys.path.insert(0, '.')
from softgroup_b200 import harness, ops, profiler, synth  # noqa: E402
from softgroup_b200.configs import model_cfg  # noqa: E402
from softgroup_b200.model import SoftGroup  # noqa: E402
from softgroup_b200.ops import _lib  # noqa: E402

L = _lib.lib()
L.sgb_test_set_tc_tuning.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
L.sgb_test_set_tc_tuning.restype = None

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


def run(tag, pf, nb, sk, steps=8):
    L.sgb_test_set_tc_tuning(pf, nb, sk)
    with torch.no_grad():
        for _ in range(3):
            out = step()
        torch.cuda.synchronize()
        evs = []
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = step()
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in evs)
        profiler.reset()
        for _ in range(3):
            flush.zero_()
            profiler.enable()
            step()
            profiler.disable()
        s = profiler.summary()['by_kernel']
    conv = {k: v['ms_per_step'] for k, v in s.items() if 'conv' in k or 'linear' in k or 'split' in k}
    print('%-28s step median %.3f min %.3f ms | %s | proposals %d' %
          (tag, ms[len(ms) // 2], ms[0], ' '.join('%s=%.3f' % kv for kv in sorted(conv.items())),
           out['proposals_offset'].numel() - 1), flush=True)
    return out


for name in ('sgb_test_set_tc_gather', 'sgb_test_set_tc_split_policy', 'sgb_test_set_bfs_mode'):
    getattr(L, name).argtypes = [ctypes.c_int]
    getattr(L, name).restype = None


def knobs(gather=0, split_policy=0, bfs_mode=0):
    L.sgb_test_set_tc_gather(gather)
    L.sgb_test_set_tc_split_policy(split_policy)
    L.sgb_test_set_bfs_mode(bfs_mode)


knobs()
ref = run('base (validated defaults)', 0, 3, 8)
for tag, kw in [('bfs frontier labelling', dict(bfs_mode=1)), ('wave-aware split-K', dict(split_policy=1)),
                ('cooperative gather', dict(gather=1)), ('all three', dict(gather=1, split_policy=1, bfs_mode=1))]:
    knobs(**kw)
    out = run(tag, 0, 3, 8)
    for k in ('proposals_idx', 'proposals_offset'):
        assert torch.equal(out[k], ref[k]), (tag, k)  # grouping is bit-exact by construction
    # the instance head's scores go through the backbone features and the tiny U-Net: they see every conv variant
    d = (out['device_instances']['score'] - ref['device_instances']['score']).abs().max().item()
    print('   max |d instance score| = %.3g (scale %.3g)' % (d, ref['device_instances']['score'].abs().max().item()))
knobs()
