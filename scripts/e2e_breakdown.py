"""Where the end-to-end time goes (host side): cProfile of run_scan."""
import cProfile, pstats, sys, time
import torch
sys.path.insert(0, '.')
from softgroup_b200 import harness, synth
from softgroup_b200.configs import model_cfg
from softgroup_b200.model import SoftGroup
torch.manual_seed(0)
model = SoftGroup(**model_cfg('scannet')).cuda().eval()
scan = synth.make_scan('c2_scannet', seed=0)
hb = harness.to_host_batch(scan)
inj = harness.pointwise_injection(scan)
with torch.no_grad():
    for _ in range(3):
        harness.run_scan(model, hb, inject_pointwise=inj)
    torch.cuda.synchronize()
    t = time.time()
    for _ in range(5):
        harness.run_scan(model, hb, inject_pointwise=inj)
    torch.cuda.synchronize()
    print('e2e ms', (time.time() - t) / 5 * 1e3)
    t = time.time()
    for _ in range(5):
        harness.run_scan(model, hb, inject_pointwise=inj, device_only=True)
    torch.cuda.synchronize()
    print('device_only ms', (time.time() - t) / 5 * 1e3)
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(3):
        harness.run_scan(model, hb, inject_pointwise=inj)
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
