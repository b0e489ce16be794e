"""CPU: the module tree / state_dict contract. When the reference tree is present (build container only),
the UNMODIFIED reference SoftGroup class is instantiated on top of this repo's spconv/ops shims and its
state_dict keys and shapes must equal ours one to one (checkpoints load unchanged, softgroup/util/utils.py:111-145)."""
import os
import sys

import pytest
import torch

from softgroup_b200.configs import model_cfg
from softgroup_b200.model import SoftGroup

REF = '/root/reference'


def test_state_dict_names():
    m = SoftGroup(**model_cfg('scannet'))
    sd = m.state_dict()
    for k in ['input_conv.0.weight', 'unet.blocks.block0.conv_branch.2.weight', 'unet.conv.2.weight',
              'unet.deconv.2.weight', 'unet.blocks_tail.block0.i_branch.0.weight', 'unet.u.u.u.u.u.u.blocks.block1.conv_branch.5.weight',
              'tiny_unet.blocks.block0.conv_branch.0.running_mean', 'semantic_linear.3.weight', 'offset_linear.1.running_var',
              'cls_linear.weight', 'mask_linear.2.bias', 'iou_score_linear.bias', 'output_layer.0.weight']:
        assert k in sd, k
    assert tuple(sd['input_conv.0.weight'].shape) == (32, 3, 3, 3, 6)
    assert tuple(sd['unet.conv.2.weight'].shape) == (64, 2, 2, 2, 32)
    assert tuple(sd['unet.deconv.2.weight'].shape) == (32, 2, 2, 2, 64)
    assert tuple(sd['unet.blocks_tail.block0.i_branch.0.weight'].shape) == (32, 1, 1, 1, 64)
    n_subm = sum(1 for k, v in sd.items() if k.endswith('weight') and v.dim() == 5 and v.shape[1] == 3)
    assert n_subm == 53 + 12  # backbone + tiny U-Net


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree only exists in the build container')
def test_reference_class_on_our_shims_has_identical_state_dict():
    import importlib
    import types
    import softgroup_b200
    softgroup_b200.install_as_reference_backends()
    # load the reference's model package without importing softgroup/__init__ side effects
    pkg = types.ModuleType('softgroup')
    pkg.__path__ = [os.path.join(REF, 'softgroup')]
    sys.modules.setdefault('softgroup', pkg)
    util = types.ModuleType('softgroup.util')
    from softgroup_b200 import util as our_util
    for n in ('cuda_cast', 'force_fp32', 'rle_decode', 'rle_encode'):
        setattr(util, n, getattr(our_util, n))
    sys.modules.setdefault('softgroup.util', util)
    ref_model = importlib.import_module('softgroup.model.softgroup')
    cfg = model_cfg('scannet')
    ref = ref_model.SoftGroup(**cfg)
    ours = SoftGroup(**cfg)
    a, b = ref.state_dict(), ours.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k].shape == b[k].shape, k
    ours.load_state_dict(a)  # strict


def test_model_survives_deepcopy_and_pickle():
    """The per-thread scratch (threading.local) and the compiled plans (ctypes arrays, raw pointers) stay out of the module's
    state: copy.deepcopy and torch.save(module) work and the copies rebuild them on demand."""
    import copy
    import io
    from softgroup_b200.configs import model_cfg
    from softgroup_b200.model import SoftGroup
    m = SoftGroup(**model_cfg('scannet', channels=16, num_blocks=3))
    m2 = copy.deepcopy(m)
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    m3 = torch.load(buf, weights_only=False)
    for c in (m2, m3):
        assert c._plans == {} and hasattr(c, '_tls')
        assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), c.state_dict().values()))
