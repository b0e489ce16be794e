"""get_instances host logic on CPU tensors (pure torch + the library's host RLE formatter, no GPU): the sparse
procedure of softgroup_b200/model/softgroup.py against a dense restatement of the reference's
softgroup/model/softgroup.py:537-604, with and without lvl_fusion (voxel -> point expansion through v2p_map)."""
import numpy as np
import pytest
import torch

from softgroup_b200.configs import model_cfg
from softgroup_b200.model import SoftGroup
from softgroup_b200.model.softgroup import expand_voxel_entries
from softgroup_b200.util.rle import rle_encode


def reference_get_instances(scan_id, proposals_idx, semantic_scores, cls_scores, iou_scores, mask_scores, instance_classes,
                            sem2ins_classes, cls_score_thr, mask_score_thr, min_npoint, v2p_map=None, lvl_fusion=False):
    """Dense restatement of the reference procedure (per-class dense [nInst, rows] masks)."""
    num_instances = cls_scores.shape[0]
    num_rows = semantic_scores.shape[0]
    cls_sm = torch.from_numpy(cls_scores).softmax(1).numpy()
    semantic_pred = semantic_scores.argmax(1)
    out = []
    for i in range(instance_classes):
        if i in sem2ins_classes:
            mask = (semantic_pred == i)[None, :].astype(np.int32)
            if lvl_fusion:
                mask = mask[:, v2p_map]
            out.append((i + 1, np.float32(1.), mask[0]))
            continue
        score = cls_sm[:, i] * np.clip(iou_scores[:, i], 0, 1)
        mask = np.zeros((num_instances, num_rows), np.int32)
        sel = mask_scores[:, i] > mask_score_thr
        mask[proposals_idx[sel, 0], proposals_idx[sel, 1]] = 1
        keep = cls_sm[:, i] > cls_score_thr
        score, mask = score[keep], mask[keep]
        if lvl_fusion:
            mask = mask[:, v2p_map]
        keep2 = mask.sum(1) >= min_npoint
        for s, m in zip(score[keep2], mask[keep2]):
            out.append((i + 1, s, m))
    return [dict(scan_id=scan_id, label_id=l, conf=c, pred_mask=rle_encode(m)) for l, c, m in out]


def _case(seed, rows, n_prop, n_classes_sem, n_inst):
    rng = np.random.RandomState(seed)
    lens = rng.randint(5, 60, n_prop)
    pidx = []
    for p, ln in enumerate(lens):
        members = rng.choice(rows, ln, replace=False)  # BFS order = arbitrary order
        pidx.append(np.stack([np.full(ln, p), members], 1))
    pidx = np.concatenate(pidx).astype(np.int32)
    sem = rng.randn(rows, n_classes_sem).astype(np.float32)
    cls = (rng.randn(n_prop, n_inst + 1) * 2).astype(np.float32)
    iou = rng.uniform(-0.2, 1.2, (n_prop, n_inst + 1)).astype(np.float32)
    msk = rng.randn(pidx.shape[0], n_inst + 1).astype(np.float32)
    return pidx, sem, cls, iou, msk


def _model(name, **test_cfg):
    cfg = model_cfg(name, channels=16, num_blocks=2, test_cfg=test_cfg)
    return SoftGroup(**cfg).eval()


@pytest.mark.parametrize('name,seed', [('scannet', 0), ('scannet', 1), ('s3dis', 2)])
def test_get_instances_matches_dense_procedure(host_instance_ops, name, seed):
    model = _model(name, min_npoint=8, cls_score_thr=0.02, mask_score_thr=-0.5)
    rows = 400
    pidx, sem, cls, iou, msk = _case(seed, rows, 14, model.semantic_classes, model.instance_classes)
    got = model.get_instances('scan0', torch.from_numpy(pidx), torch.from_numpy(sem), torch.from_numpy(cls),
                              torch.from_numpy(iou), torch.from_numpy(msk))
    want = reference_get_instances('scan0', pidx, sem, cls, iou, msk, model.instance_classes, model.sem2ins_classes, 0.02,
                                   -0.5, 8)
    assert len(got) == len(want) > 0
    for g, w in zip(got, want):
        assert g['label_id'] == w['label_id'] and g['pred_mask'] == w['pred_mask'] and g['scan_id'] == 'scan0'
        assert np.float32(g['conf']) == np.float32(w['conf'])


@pytest.mark.parametrize('name,seed', [('scannet', 3), ('s3dis', 4)])
def test_get_instances_lvl_fusion_matches_dense_procedure(host_instance_ops, name, seed):
    model = _model(name, min_npoint=12, cls_score_thr=0.02, mask_score_thr=-0.5)
    n_vox, n_pts = 300, 1000
    rng = np.random.RandomState(100 + seed)
    v2p = rng.randint(0, n_vox, n_pts).astype(np.int32)  # some voxels hold several points, some none
    pidx, sem, cls, iou, msk = _case(seed, n_vox, 12, model.semantic_classes, model.instance_classes)
    got = model.get_instances('s', torch.from_numpy(pidx), torch.from_numpy(sem), torch.from_numpy(cls),
                              torch.from_numpy(iou), torch.from_numpy(msk), v2p_map=torch.from_numpy(v2p),
                              lvl_fusion=True)
    want = reference_get_instances('s', pidx, sem, cls, iou, msk, model.instance_classes, model.sem2ins_classes, 0.02,
                                   -0.5, 12, v2p_map=v2p.astype(np.int64), lvl_fusion=True)
    assert len(got) == len(want) > 0
    for g, w in zip(got, want):
        assert g['label_id'] == w['label_id'] and g['pred_mask'] == w['pred_mask']
        assert g['pred_mask']['length'] == n_pts
        assert np.float32(g['conf']) == np.float32(w['conf'])


def test_expand_voxel_entries():
    v2p = torch.tensor([2, 0, 2, 1, 2, 0], dtype=torch.int32)  # voxel 3 is empty
    entries = torch.tensor([[0, 2], [0, 3], [1, 0], [1, 2]], dtype=torch.int32)
    out, row_of = expand_voxel_entries(entries, v2p, 4)
    assert out.dtype == torch.int32
    assert out.tolist() == [[0, 0], [0, 2], [0, 4], [1, 1], [1, 5], [1, 0], [1, 2], [1, 4]]
    assert row_of.tolist() == [0, 0, 0, 2, 2, 3, 3, 3]
    out, row_of = expand_voxel_entries(entries[:0], v2p, 4)
    assert out.shape == (0, 2) and row_of.numel() == 0
