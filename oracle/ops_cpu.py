"""oracle/ops_cpu.py -- the `softgroup.ops` function surface (softgroup/ops/functions.py) on CPU torch tensors, every
function a thin wrapper over the oracle (oracle/sg_oracle.c). TEST INFRASTRUCTURE ONLY: it lets the UNMODIFIED
reference model code run in the build container without a GPU so that whole-model golden vectors can be generated
(tests/golden/make_forward_golden.py). Signatures and return layouts follow the reference wrappers:
ball_query :7-11, octree_ball_query :14-44, voxelization_idx :168-197, voxelization :200-234,
ballquery_batch_p :237-275, bfs_cluster :278-308, global_avg_pool :311-348, sec_mean/min/max :351-438,
get_mask_iou_on_cluster/_on_pred :47-125, get_mask_label :128-165."""
import sys

import numpy as np
import torch

import oracle as _o


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def _n(t):
    return t.detach().cpu().numpy()


def voxelization_idx(coords, batchsize, mode=4):
    oc, im, om = _o.voxelization_idx(_n(coords), batchsize, mode)
    return _t(oc), _t(im), _t(om)


def voxelization(feats, map_rule, mode=4):
    return _t(_o.voxelization(_n(feats), _n(map_rule), mode))


def ballquery_batch_p(coords, batch_idxs, batch_offsets, radius, meanActive):
    idx, sl = _o.ballquery_batch_p(_n(coords), _n(batch_idxs), _n(batch_offsets), radius)
    return _t(idx), _t(sl)


def octree_ball_query(coords, mean_active, radius):
    idx, sl = _o.octree_ball_query(_n(coords), mean_active, radius)
    return _t(idx), _t(sl)


def ball_query(coords, batch_idxs, batch_offsets, radius, mean_active, with_octree=False):
    if with_octree:
        return octree_ball_query(coords, mean_active, radius)
    return ballquery_batch_p(coords, batch_idxs, batch_offsets, radius, mean_active)


def bfs_cluster(cluster_numpoint_mean, ball_query_idxs, start_len, threshold, class_id):
    a, b = _o.bfs_cluster(_n(cluster_numpoint_mean), _n(ball_query_idxs), _n(start_len), threshold, class_id)
    return _t(a), _t(b)


def global_avg_pool(feats, proposals_offset):
    return _t(_o.global_avg_pool(_n(feats), _n(proposals_offset)))


def sec_mean(inp, offsets):
    return _t(_o.sec_mean(_n(inp), _n(offsets)))


def sec_min(inp, offsets):
    return _t(_o.sec_min(_n(inp), _n(offsets)))


def sec_max(inp, offsets):
    return _t(_o.sec_max(_n(inp), _n(offsets)))


def get_mask_iou_on_cluster(proposals_idx, proposals_offset, instance_labels, instance_pointnum):
    return _t(_o.get_mask_iou_on_cluster(_n(proposals_idx), _n(proposals_offset), _n(instance_labels),
                                         _n(instance_pointnum)))


def get_mask_iou_on_pred(proposals_idx, proposals_offset, instance_labels, instance_pointnum, mask_scores_sigmoid):
    return _t(_o.get_mask_iou_on_pred(_n(proposals_idx), _n(proposals_offset), _n(instance_labels),
                                      _n(instance_pointnum), _n(mask_scores_sigmoid)))


def get_mask_label(proposals_idx, proposals_offset, instance_labels, instance_cls, instance_pointnum, proposals_iou,
                   iou_thr):
    return _t(_o.get_mask_label(_n(proposals_idx), _n(proposals_offset), _n(instance_labels), _n(instance_cls),
                                _n(instance_pointnum), _n(proposals_iou), iou_thr))


def install(name='softgroup.ops'):
    """Register under the name the reference imports (`from ..ops import ...`, softgroup/model/softgroup.py:11-13)."""
    sys.modules[name] = sys.modules[__name__]
    return sys.modules[__name__]


# ---------------------------------------------------------------------------------------------------------------------
# Dense stand-ins for softgroup_b200.ops.instances (the GPU bitmap path of get_instances): exactly the reference's
# dense-mask steps, softgroup/model/softgroup.py:551-580 and softgroup/util/rle.py:5-19. CPU tests only.
# ---------------------------------------------------------------------------------------------------------------------
def instance_point_counts(proposals_idx, mask_scores, n_classes, thr, n_proposals):
    pidx, ms = _n(proposals_idx), _n(mask_scores)
    out = np.zeros((n_proposals, n_classes), np.int32)
    for i in range(n_classes):
        on = ms[:, i] > thr
        np.add.at(out[:, i], pidx[on, 0], 1)  # (proposal, point) pairs are unique: == mask_pred.sum(1) (:563-565)
    return _t(out)


def instance_bitmaps(proposals_idx, mask_scores, n_classes, thr, keep, n_points):
    """'bitmaps' are dense uint8 masks [n_inst, n_points] here (:553-555 `mask_pred[proposals_idx[...]] = 1`)."""
    pidx, ms, kp_ = _n(proposals_idx), _n(mask_scores), _n(keep)
    kc, kp = np.nonzero(kp_.T)
    masks = np.zeros((kc.size, int(n_points)), np.uint8)
    for k, (i, p) in enumerate(zip(kc, kp)):
        sel = (pidx[:, 0] == p) & (ms[:, i] > thr)
        masks[k, pidx[sel, 1]] = 1
    return _t(masks), _t(kc.astype(np.int64)), _t(kp.astype(np.int64))


def bitmaps_to_rle(masks, n_points):
    out = []
    for m in _n(masks):
        m = np.concatenate([[0], m, [0]])  # rle.py:13-17
        runs = np.where(m[1:] != m[:-1])[0] + 1
        runs[1::2] -= runs[::2]
        out.append(dict(length=int(n_points), counts=' '.join(str(x) for x in runs)))
    return out
