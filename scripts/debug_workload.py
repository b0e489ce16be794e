"""Debug helper: one device step of a bench workload with the grouping internals printed.
Usage: python scripts/debug_workload.py c3|c5"""
import sys
import traceback

import torch

sys.path.insert(0, '.')
import bench  # noqa: E402
from softgroup_b200 import harness, ops, synth  # noqa: E402
from softgroup_b200.configs import model_cfg  # noqa: E402
from softgroup_b200.model import SoftGroup  # noqa: E402
from softgroup_b200.model import softgroup as sg  # noqa: E402

wl = bench.WORKLOADS[sys.argv[1]]
cfg = model_cfg(wl['cfg'])
sc = synth.make_scan(wl['shape'], seed=0, n_points=wl['n'])
if wl.get('intensity_only'):
    sc['feats'] = sc['feats'][:, :1].copy()
base = sc
if wl.get('x4'):
    sc = synth.to_x4_split(sc)
torch.manual_seed(0)
model = SoftGroup(**cfg).cuda().eval()
hb = harness.to_host_batch(sc, pin=False)
inj = harness.pointwise_injection(base, sigma=wl['sigma'], seed=0, fragments=wl.get('fragments', 1), confusion=wl.get('confusion', 0.0))
dev = harness.device_batch(hb)
print('points', hb['coords'].shape, 'batch_size', hb['batch_size'], 'inj', inj[0].shape, inj[1].shape, flush=True)

orig_bfs = sg.bfs_cluster_segments


def bfs(ni, sl, thr, **kw):
    print('  bfs: nodes', sl.size(0), 'list entries', ni.numel(), 'sum len', int(sl[:, 1].long().sum()), 'max start', int(sl[:, 0].max()),
          'thr', thr, flush=True)
    pidx, poff = orig_bfs(ni, sl, thr, **kw)
    print('  bfs ->', tuple(pidx.shape), tuple(poff.shape), 'max cluster id', int(pidx[:, 0].max()) if pidx.numel() else None,
          'max node', int(pidx[:, 1].max()) if pidx.numel() else None, flush=True)
    if pidx.numel() and int(pidx[:, 0].max()) >= poff.numel() - 1 and not kw.get('node_seg'):
        import numpy as np
        import oracle
        g, o = pidx.cpu().numpy(), poff.cpu().numpy()
        mean = np.full(20, -1, np.float32)
        wi, wo = oracle.bfs_cluster(mean, ni.cpu().numpy(), sl.cpu().numpy(), float(thr), 0)
        print('  ORACLE', wi.shape, wo.shape, 'offsets equal', np.array_equal(o, wo), flush=True)
        bad = np.nonzero((g != wi).any(1))[0]
        print('  mismatching rows', bad.size, 'first', bad[:10], flush=True)
        for r in bad[:6]:
            c = int(np.searchsorted(wo, r, side='right') - 1)
            print('   row', r, 'cluster', c, 'range', wo[c], wo[c + 1], 'got', g[r], 'want', wi[r], flush=True)
        lens = sl.cpu().numpy()[:, 1]
        print('  list lengths: max', lens.max(), 'hist >1000:', int((lens > 1000).sum()), flush=True)
    return pidx, poff


sg.bfs_cluster_segments = bfs
if hasattr(sg, 'group_entries'):
    orig_ge = sg.group_entries

    def ge(*a, **k):
        r = orig_ge(*a, **k)
        print('  group_entries total/per class', r[4].tolist(), 'seg_offsets', r[3].tolist()[:12], flush=True)
        return r
    sg.group_entries = ge
with torch.no_grad():
    try:
        vc, v2p, p2v = ops.voxelization_idx(dev['coords'], dev['batch_size'])
        d = {k: v for k, v in dev.items() if k not in ('coords', 'voxel_coords', 'v2p_map', 'p2v_map')}
        print('scores argmax histogram', torch.bincount(inj[0].argmax(1)).tolist(), flush=True)
        out = model.forward_test(device_only=True, inject_pointwise=inj, voxel_coords=vc, v2p_map=v2p, p2v_map=p2v, **d)
        print('proposals', out['proposals_offset'].numel() - 1, out['proposals_idx'].shape)
    except Exception:
        traceback.print_exc()
