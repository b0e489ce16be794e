import ctypes, sys, numpy as np, torch
sys.path.insert(0, '.')
from softgroup_b200 import spconv
from softgroup_b200.ops import _lib
L = _lib.lib()
L.sgb_test_set_tc_debug.argtypes = [ctypes.c_void_p]
C = int(sys.argv[1]) if len(sys.argv) > 1 else 32
M = int(sys.argv[2]) if len(sys.argv) > 2 else 137000
rng = np.random.RandomState(0)
# surface-like sparse set: random voxels on a few planes
pts = np.unique(np.stack([np.zeros(M, int), rng.randint(0, 300, M), rng.randint(0, 300, M), rng.randint(0, 3, M)], 1), axis=0).astype(np.int32)
pts = pts[rng.permutation(len(pts))]
idx = torch.from_numpy(pts).cuda()
feats = torch.randn(idx.size(0), C, device='cuda')
conv = spconv.SubMConv3d(C, C, 3, padding=1, bias=False, indice_key='k').cuda()
x = spconv.SparseConvTensor(feats, idx, (400, 400, 128), 1)
with torch.no_grad():
    conv(x); conv(x)
    torch.cuda.synchronize()
    dbg = torch.zeros(64 * 8, dtype=torch.int64, device='cuda')
    L.sgb_test_set_tc_debug(ctypes.c_void_p(dbg.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); conv(x); e1.record()
    torch.cuda.synchronize()
    L.sgb_test_set_tc_debug(None)
print('rows', idx.size(0), 'C', C, 'conv ms', e0.elapsed_time(e1))
d = dbg.cpu().numpy().reshape(64, 8)
t0 = d[d > 0].min()
print('iter: P.wait_start P.wait_done P.sts_done P.fence_done | M.start M.bfull M.afull M.issued   (cycles from start)')
for i in range(28):
    print(i, ' '.join('%7d' % (v - t0 if v > 0 else -1) for v in d[i]))
