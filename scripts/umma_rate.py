import ctypes, sys, torch
sys.path.insert(0, '.')
from softgroup_b200.ops import _lib
L = _lib.lib()
L.sgb_test_umma_rate.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
out = torch.zeros(1, dtype=torch.int64, device='cuda')
names = {100: '[2nt@D | nt@D+nt] overlapping (2 MMA/k-step)', 101: '3 x nt into one accumulator (3 MMA/k-step)',
         102: '[2nt@D | nt@D2] disjoint (2 MMA/k-step)'}
for mode in (100, 101, 102):
    for N in (32, 48, 64, 96, 128):
        for per in (2, 96):
            reps = 960
            for _ in range(2):
                rc = L.sgb_test_umma_rate(N, reps, per, ctypes.c_void_p(out.data_ptr()), None, mode)
                assert rc == 0, L.sgb_last_error()
                torch.cuda.synchronize()
            print('%-48s nt=%3d k-steps/commit=%3d: %.1f cycles per k-step' % (names[mode], N, per, out.item() / reps))
