"""In-kernel clock64 timeline of one CTA of the tensor-core conv. Usage: tc_timeline.py C M [skip_mask] [ksplit]"""
import ctypes, sys, numpy as np, torch
sys.path.insert(0, '.')
from softgroup_b200 import spconv
from softgroup_b200.ops import _lib
L = _lib.lib()
L.sgb_test_set_tc_debug.argtypes = [ctypes.c_void_p]
L.sgb_test_set_tc_tuning.argtypes = [ctypes.c_int] * 3
L.sgb_test_set_tc_tuning.restype = None
L.sgb_test_set_tc_skip.argtypes = [ctypes.c_int]
L.sgb_test_set_tc_skip.restype = None
C = int(sys.argv[1]) if len(sys.argv) > 1 else 32
M = int(sys.argv[2]) if len(sys.argv) > 2 else 137000
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
L.sgb_test_set_tc_tuning(0, 3, int(sys.argv[4]) if len(sys.argv) > 4 else 1)
rng = np.random.RandomState(0)
side = max(8, int((M / 2.0) ** 0.5))
pts = np.unique(np.stack([np.zeros(M, int), rng.randint(0, side, M), rng.randint(0, side, M), rng.randint(0, 3, M)], 1), axis=0).astype(np.int32)
pts = pts[rng.permutation(len(pts))]
idx = torch.from_numpy(pts).cuda()
feats = torch.randn(idx.size(0), C, device='cuda')
conv = spconv.SubMConv3d(C, C, 3, padding=1, bias=False, indice_key='k').cuda()
x = spconv.SparseConvTensor(feats, idx, (side + 8, side + 8, 128), 1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
with torch.no_grad():
    conv(x); conv(x)
    L.sgb_test_set_tc_skip(skip)
    conv(x)
    torch.cuda.synchronize()
    dbg = torch.zeros(65 * 8, dtype=torch.int64, device='cuda')
    L.sgb_test_set_tc_debug(ctypes.c_void_p(dbg.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); conv(x); e1.record()
    torch.cuda.synchronize()
    L.sgb_test_set_tc_debug(None)
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); conv(x); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    L.sgb_test_set_tc_skip(0)
print('rows', idx.size(0), 'C', C, 'skip', skip, 'conv+act_split ms (dbg run)', round(e0.elapsed_time(e1), 4), 'plain runs', [round(t, 4) for t in ts])
d = dbg.cpu().numpy().reshape(65, 8)
t0 = d[64, 0]
print('CTA marks: entry=0 tmem_alloc=%d map_loaded=%d roles_start=%d mainloop_done(tid0)=%d acc_done=%d epilogue_done=%d exit=%d' % tuple(int(v - t0) for v in d[64, 1:]))
print('row i: producer iteration i: entry free_ok fenced arrived+reloaded | row P: MMA pair P: start bfull afull committed')
for i in range(28 if skip & 32 else 0):
    print(i, ' '.join('%7d' % (v - t0 if v > 0 else -1) for v in d[i]))
