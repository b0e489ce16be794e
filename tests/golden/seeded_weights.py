"""Deterministic, platform-independent weights for golden vectors: every parameter / buffer of a SoftGroup-shaped
module tree is filled from numpy's legacy RandomState seeded by (seed, crc32 of the tensor's name), so the generator
(reference model, CPU) and the tests (this repo's model, GPU) build bit-identical state_dicts without shipping them."""
import zlib

import numpy as np
import torch

# The fixture's configuration (tests/golden/ref_forward_c1.npz): small model; grouping thresholds fitted to a 2 000-point
# scan (5 cm point spacing, objects of a few hundred points) -- the yaml values (radius 0.04, thresholds relative to
# ScanNet class means of thousands of points) would give no proposal there.
CFG_OVERRIDES = dict(channels=16, num_blocks=4, test_cfg=dict(min_npoint=20),
                     grouping_cfg=dict(radius=0.25, class_numpoint_mean=[-1.] * 20, npoint_thr=30))
SCAN = dict(shape='c1_plumbing', seed=0)
WEIGHT_SEED = 1234
CALIBRATED_KEYS = ('semantic_linear.3.weight', 'semantic_linear.3.bias', 'offset_linear.3.weight', 'offset_linear.3.bias')


def _rng(seed, name):
    return np.random.RandomState((seed * 1000003 + zlib.crc32(name.encode())) % (2**32))


def fill_seeded(model, seed):
    with torch.no_grad():
        for name, p in model.named_parameters():
            r = _rng(seed, name)
            if p.dim() >= 2:  # conv [out,k,k,k,in] or linear [out,in]: He-style scale on the fan-in
                fan_in = int(np.prod(p.shape[1:]))
                v = r.randn(*p.shape) * np.sqrt(2.0 / fan_in)
            elif name.endswith('weight'):  # BatchNorm1d weight
                v = r.uniform(0.8, 1.2, p.shape)
            else:  # biases
                v = r.randn(*p.shape) * 0.05
            p.copy_(torch.from_numpy(v.astype(np.float32)))
        for name, b in model.named_buffers():
            r = _rng(seed, name)
            if name.endswith('running_mean'):
                b.copy_(torch.from_numpy((r.randn(*b.shape) * 0.1).astype(np.float32)))
            elif name.endswith('running_var'):
                b.copy_(torch.from_numpy(r.uniform(0.5, 1.5, b.shape).astype(np.float32)))
    return model


def weights_digest(model):
    """crc32 over all parameters and float buffers in state_dict order (a cheap identity check for the fixture)."""
    crc = 0
    for k, v in model.state_dict().items():
        if v.dtype.is_floating_point:
            crc = zlib.crc32(v.detach().cpu().numpy().astype(np.float32).tobytes(), crc)
    return crc


def load_calibrated(model, flat):
    """Put the fixture's closed-form head layers (stored flat, CALIBRATED_KEYS order) into `model`."""
    sd = model.state_dict()
    pos = 0
    with torch.no_grad():
        for k in CALIBRATED_KEYS:
            n = sd[k].numel()
            sd[k].copy_(torch.from_numpy(np.asarray(flat[pos:pos + n], np.float32)).view_as(sd[k]))
            pos += n
    assert pos == len(flat)
    return model
