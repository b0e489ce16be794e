"""CPU: anchor the sparse-conv restatement (oracle/spconv_oracle.py) on dense torch convolutions.
spconv itself is absent from the reference tree ("parity unpinned"), so the dense cross-check
IS the spec: weight [out,k0,k1,k2,in], cross-correlation, odd dims drop their max plane at a stride-2 level,
the inverse conv restores the paired conv's sites."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import spconv_oracle as so


def _sparse_case(seed, shape=(11, 12, 9), B=2, Cin=5, density=0.15):
    rng = np.random.RandomState(seed)
    idx = []
    for b in range(B):
        occ = np.argwhere(rng.rand(*shape) < density)
        idx.append(np.concatenate([np.full((len(occ), 1), b), occ], 1))
    idx = np.concatenate(idx, 0).astype(np.int32)
    idx = idx[rng.permutation(len(idx))]
    feats = rng.randn(len(idx), Cin).astype(np.float32)
    return idx, feats, shape, B


def _dense(idx, feats, shape, B):
    C = feats.shape[1]
    d = torch.zeros((B, C) + tuple(shape), dtype=torch.float64)
    d[idx[:, 0], :, idx[:, 1], idx[:, 2], idx[:, 3]] = torch.from_numpy(feats).double()
    return d


@pytest.mark.parametrize('seed', [0, 1])
def test_subm_matches_dense_conv3d(seed):
    idx, feats, shape, B = _sparse_case(seed)
    rng = np.random.RandomState(seed + 10)
    W = rng.randn(7, 3, 3, 3, feats.shape[1]).astype(np.float32)
    out = so.subm_conv3d(feats, idx, W, acc64=True)
    dense = F.conv3d(_dense(idx, feats, shape, B), torch.from_numpy(W).double().permute(0, 4, 1, 2, 3), padding=1)
    want = dense[idx[:, 0], :, idx[:, 1], idx[:, 2], idx[:, 3]].numpy()
    np.testing.assert_allclose(out, want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('shape', [(11, 12, 9), (8, 8, 8), (13, 7, 5)])
def test_down_and_inverse_match_dense(shape):
    idx, feats, shape, B = _sparse_case(3, shape=shape)
    rng = np.random.RandomState(4)
    Cin = feats.shape[1]
    W = rng.randn(6, 2, 2, 2, Cin).astype(np.float32)
    out, out_idx, inv, out_shape = so.sparse_conv3d_k2s2(feats, idx, shape, W, acc64=True)
    assert out_shape == [s // 2 for s in shape]
    dense = F.conv3d(_dense(idx, feats, shape, B), torch.from_numpy(W).double().permute(0, 4, 1, 2, 3), stride=2)
    assert list(dense.shape[2:]) == out_shape
    want = dense[out_idx[:, 0], :, out_idx[:, 1], out_idx[:, 2], out_idx[:, 3]].numpy()
    np.testing.assert_allclose(out, want, rtol=1e-5, atol=1e-5)
    # every non-zero dense output site is an output voxel (and vice versa for generic weights)
    nz = (dense.abs().sum(1) > 0).nonzero().numpy()
    assert set(map(tuple, nz.tolist())) == set(map(tuple, out_idx.tolist()))
    # first-occurrence numbering
    par = idx.copy()
    par[:, 1:] >>= 1
    valid = np.all(par[:, 1:] < np.array(out_shape), 1)
    seen, order = set(), []
    for r in par[valid]:
        t = tuple(r)
        if t not in seen:
            seen.add(t)
            order.append(t)
    assert order == list(map(tuple, out_idx.tolist()))
    # inverse conv: out[child] = W[:, o, :] . in[parent]; dropped children (max plane of an odd dim) get 0
    Wi = rng.randn(Cin, 2, 2, 2, 6).astype(np.float32)
    back = so.inverse_conv3d_k2(out, inv, Wi, acc64=True)
    dT = F.conv_transpose3d(_dense(out_idx, out, out_shape, B), torch.from_numpy(Wi).double().permute(4, 0, 1, 2, 3),
                            stride=2)
    full = torch.zeros((B, Cin) + tuple(shape), dtype=torch.float64)
    full[:, :, :dT.shape[2], :dT.shape[3], :dT.shape[4]] = dT
    want = full[idx[:, 0], :, idx[:, 1], idx[:, 2], idx[:, 3]].numpy()
    np.testing.assert_allclose(back, want, rtol=1e-5, atol=1e-5)
    dropped = ~valid
    if dropped.any():
        assert np.all(back[dropped] == 0)
