"""N2: the GPU test-time data preparation (harness.prepare_test_batch_gpu: rotation, scale, shift,
truncation, batch column, spatial shape, point->voxel hash; S3DIS x4 split) against the reference's OWN dataset code --
CustomDataset.transform_test / __getitem__ / collate_fn (softgroup/data/custom.py:162-256) and S3DISDataset
(softgroup/data/s3dis.py:46-115), staged verbatim by oracle/build_ref.py and run on the host."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from softgroup_b200 import harness, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ref_data():
    from oracle import build_ref
    m = build_ref.import_reference_data()
    if m is None:
        pytest.skip('reference dataset modules not staged')
    return m


def _raw(seed, n, shape='c2_scannet'):
    """raw, mean-centred float32 points like prepare_data_inst.py:55-56 (before any test transform)."""
    scan = synth.make_scan(shape, seed=seed, n_points=n)
    rng = np.random.RandomState(seed)
    xyz = (rng.rand(n, 3) * np.array([7, 5, 2.6])).astype(np.float32)
    xyz -= xyz.mean(0)
    rgb = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    sem = scan['semantic_labels'].astype(np.int64)
    ins = scan['instance_labels'].astype(np.int64).copy()
    ins[ins < 0] = -100
    return xyz, rgb, sem, ins


def _dataset(cls, raw, x4_split=None):
    ds = object.__new__(cls)
    ds.voxel_cfg = SimpleNamespace(scale=50, spatial_shape=[128, 512], max_npoint=250000, min_npoint=5000)
    ds.training = False
    ds.filenames = ['scene0000_00_inst_nostuff.pth']
    ds.suffix = '_inst_nostuff.pth'
    ds.load = lambda filename: tuple(a.copy() for a in raw)
    if x4_split is not None:
        ds.x4_split = x4_split
    return ds


def _check(batch_ref, batch_gpu):
    for k in ('voxel_coords', 'v2p_map', 'p2v_map', 'coords_float', 'feats', 'batch_idxs'):
        a, b = batch_ref[k], batch_gpu[k].cpu()
        assert a.dtype == b.dtype and a.shape == b.shape, k
        assert torch.equal(a, b), k
    assert np.array_equal(np.asarray(batch_ref['spatial_shape']), np.asarray(batch_gpu['spatial_shape']))
    assert batch_ref['batch_size'] == batch_gpu['batch_size']
    assert torch.equal(batch_ref['semantic_labels'], batch_gpu['semantic_labels'].cpu())


@pytest.mark.parametrize('n', [2000, 150000])
def test_transform_collate_matches_reference_dataset(ref_data, n):
    custom, _ = ref_data
    raw = _raw(1, n)
    ds = _dataset(custom.CustomDataset, raw)
    want = ds.collate_fn([ds[0]])  # the dataloader's output (CPU numpy + CPU hash)
    got = harness.prepare_test_batch_gpu(raw[0], raw[1], raw[2], raw[3], scale=50, min_spatial_shape=128)
    assert torch.equal(want['coords'], got['coords'].cpu())
    _check(want, got)


def test_x4_split_matches_reference_s3dis_dataset(ref_data):
    _, s3dis = ref_data
    raw = _raw(2, 40003, shape='c3_s3dis')  # not a multiple of 4: the pieces have different lengths
    ds = _dataset(s3dis.S3DISDataset, raw, x4_split=True)
    want = ds.collate_fn([ds[0]])
    got = harness.prepare_test_batch_gpu(raw[0], raw[1], raw[2], raw[3], scale=50, min_spatial_shape=128, x4_split=True)
    _check(want, got)
