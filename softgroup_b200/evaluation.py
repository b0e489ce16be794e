"""GPU assignment step of the reference's instance evaluation + the numpy >= 1.24 fix.

`ScanNetEval.assign_instances_for_scan` (softgroup/evaluation/instance_eval.py:228-309) decodes every predicted RLE mask
and, for every ground-truth instance of the same label, evaluates `np.count_nonzero(np.logical_and(gts == id, mask))` --
O(nPred x nGt x N) on the host, inside a multiprocessing pool. Here the predicted masks become device bitmaps and ONE
kernel (sgb_bitmap_intersections) produces the whole [nPred, nGt] intersection matrix, the vertex counts and the void
intersections in O(set bits); the dictionaries handed to `evaluate_matches` are then built exactly like the reference
does (same keys, same order, same ints/floats), so the unmodified `evaluate_matches` / `compute_averages` /
`print_results` run on them.

Usage with the unmodified reference evaluator (tools/test.py:170-174):

    from softgroup_b200 import evaluation as sgb_eval
    sgb_eval.install_numpy_aliases()             # np.float / np.bool were removed in numpy 1.24 (instance_eval.py:46-47,80)
    scannet_eval = ScanNetEval(class_names)
    avgs = sgb_eval.evaluate(scannet_eval, pred_insts, gt_insts)   # instead of scannet_eval.evaluate(...)
"""
from copy import deepcopy

import numpy as np
import torch

from .ops import instances as inst_ops
from .util import rle_decode


def install_numpy_aliases():
    """`np.float` / `np.bool` (used at instance_eval.py:46-47,80) no longer exist in numpy >= 1.24: restore the aliases the
    reference was written against instead of editing its source."""
    for name, typ in (('float', float), ('bool', bool), ('int', int)):
        if not hasattr(np, name):
            setattr(np, name, typ)


def _mask_words(pred_mask, n_points, W):
    """RLE dict or dense array -> uint32 bitmap words [W] (bit k of word w = point 32 w + k)."""
    if isinstance(pred_mask, dict):
        pred_mask = rle_decode(pred_mask)
    m = np.not_equal(np.asarray(pred_mask), 0)
    assert m.shape[0] == n_points
    b = np.packbits(m, bitorder='little')
    out = np.zeros(W * 4, np.uint8)
    out[:b.size] = b
    return out.view(np.uint32)


def gt_instances_of(evaluator, gts):
    """instance_eval_util.get_instances (:140-152) without the O(nInst x N) rescans: unique ids + counts in one pass."""
    instances = {label: [] for label in evaluator.valid_class_labels}
    ids, counts = np.unique(gts, return_counts=True)
    for i, c in zip(ids.tolist(), counts.tolist()):
        if i == 0:
            continue
        label_id = int(i // 1000)
        if label_id in evaluator.valid_class_ids:
            instances[evaluator.id2label[label_id]].append(
                dict(instance_id=int(i), label_id=label_id, vert_count=int(c), med_dist=-1, dist_conf=0.0))
    return instances


def assign_instances_for_scan(evaluator, preds, gts, device='cuda'):
    """Same (gt2pred, pred2gt) as ScanNetEval.assign_instances_for_scan(preds, gts) (instance_eval.py:228-309)."""
    gts = np.asarray(gts)
    n_points = gts.shape[0]
    gt_instances = gt_instances_of(evaluator, gts)
    if evaluator.use_label:
        gt2pred = deepcopy(gt_instances)
        for label in gt2pred:
            for gt in gt2pred[label]:
                gt['matched_pred'] = []
    else:
        gt2pred = {}
        agnostic = []
        for _, instances in gt_instances.items():
            agnostic += deepcopy(instances)
        for gt in agnostic:
            gt['matched_pred'] = []
        gt2pred[evaluator.eval_class_labels[0]] = agnostic
    pred2gt = {label: [] for label in evaluator.eval_class_labels}
    # ---- device part: one column per ground-truth instance (in gt2pred order), -2 for void points ------------------
    col_of = {}
    for label in gt2pred:
        for gt in gt2pred[label]:
            col_of[gt['instance_id']] = len(col_of)
    n_gt = len(col_of)
    bool_void = np.logical_not(np.isin(gts // 1000, evaluator.valid_class_ids))
    uniq, inv = np.unique(gts, return_inverse=True)
    slot_of_uniq = np.array([col_of.get(int(u), -1) for u in uniq], np.int32)
    gslot = slot_of_uniq[inv.reshape(-1)]
    gslot = np.where(bool_void & (gslot < 0), -2, gslot).astype(np.int32)
    W = inst_ops.bitmap_words(n_points)
    usable = []  # preds that pass the label filter, in order
    for pred in preds:
        if evaluator.use_label and pred['label_id'] not in evaluator.id2label:
            continue
        usable.append(pred)
    if usable:
        words = np.stack([_mask_words(p['pred_mask'], n_points, W) for p in usable])
        bm = torch.from_numpy(words.view(np.int32)).to(device)
        inter, vert, void = inst_ops.bitmap_intersections(bm, torch.from_numpy(gslot).to(device), n_gt, n_points)
        inter, vert, void = inter.cpu().numpy(), vert.cpu().numpy(), void.cpu().numpy()
    # ---- host part: the reference's bookkeeping, verbatim in structure (:253-307) -----------------------------------
    num_pred_instances = 0
    for k, pred in enumerate(usable):
        if evaluator.use_label:
            label_id = pred['label_id']
            label_name = evaluator.id2label[label_id]
        else:
            label_name = evaluator.eval_class_labels[0]
        num = int(vert[k])
        if num < evaluator.min_region_sizes[0]:
            continue
        pred_instance = {}
        pred_instance['filename'] = '{}_{}'.format(pred['scan_id'], num_pred_instances)
        pred_instance['pred_id'] = num_pred_instances
        pred_instance['label_id'] = label_id if evaluator.use_label else None
        pred_instance['vert_count'] = num
        pred_instance['confidence'] = pred['conf']
        pred_instance['void_intersection'] = int(void[k])
        matched_gt = []
        for gt_num, gt_inst in enumerate(gt2pred[label_name]):
            intersection = int(inter[k, col_of[gt_inst['instance_id']]])
            if intersection > 0:
                gt_copy = gt_inst.copy()
                pred_copy = pred_instance.copy()
                gt_copy['intersection'] = intersection
                pred_copy['intersection'] = intersection
                iou = float(intersection) / (gt_copy['vert_count'] + pred_copy['vert_count'] - intersection)
                gt_copy['iou'] = iou
                pred_copy['iou'] = iou
                matched_gt.append(gt_copy)
                gt2pred[label_name][gt_num]['matched_pred'].append(pred_copy)
        pred_instance['matched_gt'] = matched_gt
        num_pred_instances += 1
        pred2gt[label_name].append(pred_instance)
    return gt2pred, pred2gt


def evaluate(evaluator, pred_list, gt_list, device='cuda', verbose=True):
    """ScanNetEval.evaluate (instance_eval.py:375-402) with the assignment on the GPU, scan after scan in this process
    (the reference forks a multiprocessing pool, which a CUDA context does not survive)."""
    install_numpy_aliases()
    matches = {}
    for i, (preds, gts) in enumerate(zip(pred_list, gt_list)):
        gt2pred, pred2gt = assign_instances_for_scan(evaluator, preds, gts, device=device)
        matches['gt_%d' % i] = dict(gt=gt2pred, pred=pred2gt)
    ap_scores, rc_scores = evaluator.evaluate_matches(matches)
    avgs = evaluator.compute_averages(ap_scores, rc_scores)
    if verbose:
        evaluator.print_results(avgs)
    return avgs
