"""CPU time (not wall time) of the host side of one end-to-end call, by function: cProfile driven by time.thread_time, so the
time spent blocked in cudaStreamSynchronize / .item() does not count. With scans in flight the host threads share the GIL:
this CPU time per scan bounds the end-to-end throughput. Usage: python scripts/host_cpu_profile.py"""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, '.')
from softgroup_b200 import harness, synth  # noqa: E402
from softgroup_b200.configs import model_cfg  # noqa: E402
from softgroup_b200.model import SoftGroup  # noqa: E402

torch.set_num_threads(8)
torch.manual_seed(0)
model = SoftGroup(**model_cfg('scannet')).cuda().eval()
scan = synth.make_scan('c2_scannet', seed=0)
hb = harness.to_host_batch(scan)
inj = harness.pointwise_injection(scan, sigma=0.03, seed=0)
with torch.no_grad():
    for _ in range(3):
        harness.run_scan(model, hb, inject_pointwise=inj)
    torch.cuda.synchronize()
    c0, w0 = time.thread_time(), time.perf_counter()
    for _ in range(10):
        harness.run_scan(model, hb, inject_pointwise=inj)
    torch.cuda.synchronize()
    print('per scan: CPU %.2f ms, wall %.2f ms' % ((time.thread_time() - c0) * 100, (time.perf_counter() - w0) * 100))
    pr = cProfile.Profile(time.thread_time)
    pr.enable()
    for _ in range(10):
        harness.run_scan(model, hb, inject_pointwise=inj)
    torch.cuda.synchronize()
    pr.disable()
st = pstats.Stats(pr, stream=sys.stdout)
st.sort_stats('cumulative').print_stats(40)
st.sort_stats('tottime').print_stats(35)
