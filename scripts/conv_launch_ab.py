"""Launch-by-launch comparison of the two convolution kernels inside the real step: two ncu launch lists of scripts/one_step.py
(SGB_CONV_SS=0 and =1) are matched by launch order. Usage: python scripts/conv_launch_ab.py tc.csv ss.csv"""
import csv
import sys


def load(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    for r in csv.DictReader(lines):
        n = r['Kernel Name']
        if 'spconv_tc_kernel' in n or 'spconv_ss_kernel' in n:
            rows.append((n.split('(')[0].split('::')[-1], r['Grid Size'], float(r['Metric Value']) / 1e3))
    return rows


a, b = load(sys.argv[1]), load(sys.argv[2])
assert len(a) == len(b), (len(a), len(b))
ta = tb = tbest = 0.0
for i, (x, y) in enumerate(zip(a, b)):
    ta += x[2]
    tb += y[2]
    tbest += min(x[2], y[2])
    print('%3d  tc grid %-14s %7.1f us | ss grid %-12s %7.1f us  %s' % (i, x[1], x[2], y[1], y[2], '<-- ss' if y[2] < x[2] else ''))
print('total tc %.1f us, ss %.1f us, best-of %.1f us' % (ta, tb, tbest))
