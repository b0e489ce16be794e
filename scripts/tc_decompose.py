"""Timing decomposition of the tensor-core conv (test hook sgb_test_set_tc_skip; results are garbage, only time counts):
which of tcgen05.st / MMA / gather / weight copy paces each U-Net level. Levels use the real scan's rulebooks."""
import ctypes
import sys

import torch

sys.path.insert(0, '.')
from softgroup_b200 import harness, ops, profiler, synth  # noqa: E402
from softgroup_b200.configs import model_cfg  # noqa: E402
from softgroup_b200.model import SoftGroup  # noqa: E402
from softgroup_b200 import spconv  # noqa: E402
from softgroup_b200.ops import voxelization  # noqa: E402
from softgroup_b200.ops import _lib  # noqa: E402

L = _lib.lib()
L.sgb_test_set_tc_skip.argtypes = [ctypes.c_int]
L.sgb_test_set_tc_skip.restype = None
torch.manual_seed(0)
model = SoftGroup(**model_cfg('scannet')).cuda().eval()
scan = synth.make_scan('c2_scannet', seed=0)
hb = harness.to_host_batch(scan)
dev = harness.device_batch(hb)
flush = torch.empty(512 << 20, dtype=torch.uint8, device='cuda')


def backbone():
    vc, v2p, p2v = ops.voxelization_idx(dev['coords'], 1)
    feats = torch.cat((dev['feats'], dev['coords_float']), 1)
    vf = voxelization(feats.contiguous(), p2v.contiguous())
    x = spconv.SparseConvTensor(vf, vc.int(), dev['spatial_shape'], 1)
    return model.forward_backbone(x, v2p)


with torch.no_grad():
    for mask, tag in [(0, 'normal'), (1, 'no tcgen05.st'), (2, 'no MMA'), (4, 'no gather'), (8, 'no weight copy'),
                      (3, 'no st, no MMA'), (15, 'skeleton (barriers only)'), (0, 'normal again')]:
        L.sgb_test_set_tc_skip(mask)
        for _ in range(2):
            backbone()
        profiler.reset()
        for _ in range(3):
            flush.zero_()
            profiler.enable()
            backbone()
            profiler.disable()
        torch.cuda.synchronize()
        groups = {}
        for name, nbytes, e0, e1 in profiler._records:
            if not name.startswith('spconv_tc_kernel'):
                continue
            key = int(nbytes() if callable(nbytes) else nbytes)
            g = groups.setdefault(key, [0, 0.0])
            g[0] += 1
            g[1] += e0.elapsed_time(e1)
        tot = sum(g[1] for g in groups.values()) / 3
        top = sorted(groups.items(), key=lambda kv: -kv[1][1])[:10]
        print('%-26s conv total %.3f ms | %s' % (tag, tot, ' '.join('%dMB x%d:%.0fus' % (k // 1000000, g[0] // 3, g[1] / g[0] * 1e3) for k, g in sorted(top))), flush=True)
    L.sgb_test_set_tc_skip(0)
