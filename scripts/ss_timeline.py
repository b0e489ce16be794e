"""Per-level A/B of the two convolution kernels (register-gather spconv_tc_kernel vs persistent shared-memory-ring
spconv_ss_kernel) on the rulebooks of the 150k-point bench scan, with the per-role wait counters of CTA 0 of the ss kernel
(development build: scripts/build_ss_timeline.sh). The two kernels run in two PROCESSES (the development library is
only loaded for the ss leg); outputs are compared through a file.
Usage: python scripts/ss_timeline.py            (driver: runs both children, prints the table)
       python scripts/ss_timeline.py child 0|1  (one kernel; writes /tmp/ss_ab_<k>.pt)"""
import os
import subprocess
import sys

import numpy as np
import torch

sys.path.insert(0, '.')

NAMES = ['gather:wait_map', 'gather:wait_empty', 'gather:wait_copies', 'iterations', 'items', 'mma:wait_map', 'mma:wait_acc_free',
         'mma:wait_full', 'weights:wait_empty', 'map:wait_free', 'epi:wait_acc', 'epi:busy', 'end_clock']


def child(which):
    from softgroup_b200.ops import _lib
    tl = os.path.join('scripts', 'experiments', 'build', 'libsgb200_tl.so')
    use_tl = which == 1 and os.path.exists(tl) and '--no-tl' not in sys.argv
    if use_tl:
        _lib.LIB_PATH = os.path.abspath(tl)
    from softgroup_b200 import ops, synth
    from softgroup_b200.spconv import core
    from softgroup_b200.ops._lib import check, ptr
    L = _lib.lib()
    dbg = None
    if use_tl:
        import ctypes
        dbg = torch.zeros(64 + 6 * 256, dtype=torch.int64, device='cuda')
        L.sgb_dev_ss_timeline.argtypes = [ctypes.c_void_p]
        L.sgb_dev_ss_timeline(ctypes.c_void_p(dbg.data_ptr()))
    scan = synth.make_scan('c2_scannet', seed=0)
    coords = torch.from_numpy(scan['coords']).cuda()
    vc, v2p, p2v = ops.voxelization_idx(coords, 1)
    idx = vc.int().contiguous()
    shape = [int(s) for s in scan['spatial_shape']]
    flush = torch.empty(512 << 20, dtype=torch.uint8, device='cuda')
    levels = []
    for lvl in range(7):
        C = 32 * (lvl + 1)
        mp = core.build_subm_map(idx)
        levels.append((lvl, C, idx.size(0), mp))
        if lvl < 6:
            idx, _, _, shape = core.build_down_map(idx, shape)
    torch.manual_seed(0)
    res = {}
    for lvl, C, M, mp in levels:
        if lvl > 4 or (os.environ.get('SS_LEVELS') and str(lvl) not in os.environ['SS_LEVELS'].split(',')):
            continue
        x = torch.randn(M, C, device='cuda')
        W = core.WeightPack((torch.randn(27, C, C, device='cuda') / (27 * C) ** 0.5).contiguous())
        out = torch.empty(M, C, device='cuda')
        pk = core.act_pack(x, C, 0, C)

        def run():
            check(L.sgb_spconv_forward_tc_ex(ptr(pk), C, M, ptr(mp), 27, M, ptr(W.tc()), C, C, None, 0, 0, None, ptr(out), C, 0,
                                             None, 0, 0, None, None, 0, 0, which, core._stream()))
        for _ in range(2):
            run()
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        nnz = int((mp >= 0).sum())
        res[lvl] = dict(us=float(np.median(ts)), out=out.cpu(), M=M, C=C, pairs=nnz / M,
                        dbg=dbg.cpu().tolist() if dbg is not None else None)
    torch.save(res, '/tmp/ss_ab_%d.pt' % which)


def main():
    for which in (0, 1):
        subprocess.check_call([sys.executable, __file__, 'child', str(which)] + [a for a in sys.argv[1:] if a.startswith('--')])
    a, b = torch.load('/tmp/ss_ab_0.pt'), torch.load('/tmp/ss_ab_1.pt')
    for lvl in sorted(a):
        ra, rb = a[lvl], b[lvl]
        diff = float((ra['out'] - rb['out']).abs().max() / ra['out'].abs().max())
        print('level %d  M %6d  C %3d  pairs/row %4.1f | tc %6.1f us  ss %6.1f us | max rel diff %.2e' %
              (lvl, ra['M'], ra['C'], ra['pairs'], ra['us'], rb['us'], diff), flush=True)
        if rb['dbg'] is not None:
            d = rb['dbg']
            its = max(d[3], 1)
            print('    CTA 0: %d items, %d iterations; cycles per iteration: ' % (d[4], d[3]) +
                  ', '.join('%s %.0f' % (NAMES[k], d[k] / its) for k in (0, 1, 2, 5, 6, 7, 8, 9, 10, 11)))
            t0 = d[13]
            print('    CTA 0 lifetime %d cycles (kernel %.0f cycles)' % (d[12] - t0, rb['us'] * 1965))
            if '--trace' in sys.argv:
                print('    slot | gather: slot free, copies issued, arrived || pair | weights issued | mma: pair ready, committed  (cycles since CTA start)')
                for g in list(range(0, 40)) + list(range(100, 120)):
                    if g < d[3]:
                        print('    %4d | %7d %7d %7d || %4d | %7d | %7d %7d' % tuple([g] + [d[64 + sl * 256 + g] - t0 for sl in (0, 1, 2)] + [g // 2] +
                                                                                  [d[64 + sl * 256 + g // 2] - t0 for sl in (5, 3, 4)]))


if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[1] == 'child':
        child(int(sys.argv[2]))
    else:
        main()
