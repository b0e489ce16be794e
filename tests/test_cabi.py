"""CPU: the C-ABI library loads, exports every symbol include/sgb200.h declares, and its host path
(voxelize_idx for DataLoader workers) matches the golden vectors. No GPU compute calls here."""
import os
import re

import numpy as np
import pytest
import torch

from softgroup_b200.ops import _lib
from softgroup_b200 import ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'sgb200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(sgb_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), n
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert L.sgb_abi_version() == 1


def test_device_probe_does_not_crash():
    assert _lib.lib().sgb_device_available() in (0, 1)


def test_error_reporting():
    L = _lib.lib()
    m, a = _lib.ctypes.c_int(), _lib.ctypes.c_int()
    h = L.sgb_voxelize_idx_cpu_begin(None, 5, 7, 4, None, _lib.ctypes.byref(m), _lib.ctypes.byref(a))
    assert h is None
    with pytest.raises(_lib.SgbError):
        _lib.check(h, 'cpu_begin')


def test_voxelization_idx_cpu_golden(golden):
    oc, im, om = ops.voxelization_idx(torch.from_numpy(golden['vox_c1_coords']), 1, 4)
    assert oc.dtype == torch.int64 and im.dtype == torch.int32 and om.dtype == torch.int32
    assert np.array_equal(oc.numpy(), golden['vox_c1_out_coords'])
    assert np.array_equal(im.numpy(), golden['vox_c1_input_map'])
    assert np.array_equal(om.numpy(), golden['vox_c1_output_map'])


@pytest.mark.parametrize('mode', [1, 2, 3, 4])
def test_voxelization_idx_cpu_ragged(golden, mode):
    oc, im, om = ops.voxelization_idx(torch.from_numpy(golden['vox_rag_coords']), 2, mode)
    assert np.array_equal(oc.numpy(), golden['vox_rag_m%d_out_coords' % mode])
    assert np.array_equal(im.numpy(), golden['vox_rag_m%d_input_map' % mode])
    assert np.array_equal(om.numpy(), golden['vox_rag_m%d_output_map' % mode])


def test_voxelization_idx_cpu_empty():
    oc, im, om = ops.voxelization_idx(torch.zeros((0, 4), dtype=torch.int64), 1, 4)
    assert oc.shape == (0, 4) and im.shape == (0, ) and om.shape == (0, 2)


def test_ops_fail_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises((AssertionError, RuntimeError)):
        ops.ballquery_batch_p(torch.zeros((4, 3)), torch.zeros(4, dtype=torch.int32),
                              torch.tensor([0, 4], dtype=torch.int32), 0.1, 10)
