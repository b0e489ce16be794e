import ctypes, sys, torch
sys.path.insert(0, '.')
from softgroup_b200.ops import _lib
L = _lib.lib()
L.sgb_test_umma_rate.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
out = torch.zeros(1, dtype=torch.int64, device='cuda')
for ts in (0, -1, -2, -3, 2, 3):
  for N in (32, 64):
    for per in (12, 96):
        reps = 960
        rc = L.sgb_test_umma_rate(N, reps, per, ctypes.c_void_p(out.data_ptr()), None, ts)
        assert rc == 0, L.sgb_last_error()
        torch.cuda.synchronize()
        L.sgb_test_umma_rate(N, reps, per, ctypes.c_void_p(out.data_ptr()), None, ts)
        torch.cuda.synchronize()
        print('A_in_tmem=%d N=%3d per_commit=%3d: %.1f cycles/MMA' % (ts, N, per, out.item() / reps))
