"""Host-side methods of the model against the REFERENCE'S OWN methods: the unmodified reference SoftGroup class is
instantiated on this repo's shims (CPU, build container only) and its CUDA-free methods are called directly --
panoptic_fusion (softgroup.py:606-639), get_gt_instances (:641-653), merge_4_parts (:397-409),
pyramid_inverse_map (:500-507, dense matrix), get_level (:482-489)."""
import types

import numpy as np
import pytest
import torch

from softgroup_b200.configs import model_cfg
from softgroup_b200.model import SoftGroup
from softgroup_b200.util import rle_encode


def _pair(ref_model_module, name):
    if ref_model_module is None:
        pytest.skip('reference tree not mounted')
    cfg = model_cfg(name, channels=16, num_blocks=2)
    ref = ref_model_module.SoftGroup(**cfg)
    ref.eval()  # the reference's train() override does not return self
    ref.test_cfg = types.SimpleNamespace(**cfg['test_cfg'])  # the reference reads attributes (Munch in tools/test.py)
    ref.grouping_cfg = types.SimpleNamespace(**cfg['grouping_cfg'])
    return ref, SoftGroup(**cfg).eval()


@pytest.mark.parametrize('seed', range(3))
def test_panoptic_fusion(ref_model_module, seed):
    ref, ours = _pair(ref_model_module, 'kitti')
    rng = np.random.RandomState(seed)
    n = 800
    sem = rng.randint(0, 19, n)
    insts = []
    for _ in range(15):
        lo = rng.randint(0, n - 60)
        m = np.zeros(n, np.int64)
        m[lo:lo + rng.randint(5, 60)] = rng.rand() < 0.9  # overlapping runs -> the skip rule triggers
        insts.append(dict(scan_id='s', label_id=int(rng.randint(1, 9)), conf=float(rng.rand()), pred_mask=rle_encode(m)))
    want = ref.panoptic_fusion(sem.copy(), insts)
    got = ours.panoptic_fusion(sem.copy(), insts)
    assert got.dtype == want.dtype and np.array_equal(got, want)


def test_get_gt_instances(ref_model_module):
    ref, ours = _pair(ref_model_module, 'scannet')
    rng = np.random.RandomState(0)
    sem = torch.from_numpy(rng.randint(-1, 20, 500)).long()
    sem[sem == -1] = -100
    inst = torch.from_numpy(rng.randint(-1, 12, 500)).long()
    inst[inst == -1] = -100
    want = ref.get_gt_instances(sem.clone(), inst.clone())
    got = ours.get_gt_instances(sem.clone(), inst.clone())
    assert np.array_equal(got, want)


def test_merge_4_parts(ref_model_module):
    ref, ours = _pair(ref_model_module, 's3dis')
    x = torch.arange(4 * 7 * 3, dtype=torch.float32).view(28, 3)
    assert torch.equal(ours.merge_4_parts(x), ref.merge_4_parts(x))


def test_pyramid_inverse_map_and_level(ref_model_module):
    ref, ours = _pair(ref_model_module, 'stpls3d++')
    for n in (5, 100000, 100001, 1000000, 1000001):
        assert ours.get_level(n) == ref.get_level(n)
    rng = np.random.RandomState(1)
    n_vox, n_pts = 60, 400
    l2p = torch.from_numpy(rng.randint(0, n_vox, n_pts).astype(np.int32))
    members = rng.permutation(n_vox)[:45]  # a voxel belongs to at most one cluster
    cut = [0, 10, 11, 30, 45]
    pidx = torch.tensor([[c, int(v)] for c in range(4) for v in members[cut[c]:cut[c + 1]]], dtype=torch.int32)
    poff = torch.tensor(cut, dtype=torch.int32)
    want_idx, want_off = ref.pyramid_inverse_map(pidx, poff, n_vox, l2p)
    got_idx, got_off = ours.pyramid_inverse_map(pidx, poff, n_vox, l2p)
    assert np.array_equal(got_idx.numpy(), want_idx.numpy().astype(np.int32))
    assert np.array_equal(got_off.numpy(), want_off.numpy())
