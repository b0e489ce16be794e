"""Instance masks / RLE / panoptic paste / evaluation intersections on the GPU (csrc/instances.cu through the C ABI).

Python mirror of the pieces of softgroup/model/softgroup.py:537-639 and softgroup/evaluation/instance_eval.py:228-309
that the reference runs as dense [nProposal, N] tensors + numpy loops. CUDA tensors only: there is no CPU fallback."""
import ctypes

import numpy as np
import torch

from .. import profiler
from . import _lib
from ._lib import check, ptr


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ws(nbytes, device):
    return torch.empty(int(nbytes), dtype=torch.uint8, device=device)


def bitmap_words(n_points):
    return (int(n_points) + 1 + 31) // 32


def instance_point_counts(proposals_idx, mask_scores, n_classes, thr, n_proposals):
    """npoint int32 [nP, nI]: #entries of proposal p with mask_scores[e, i] > thr (softgroup.py:553,563 `mask_pred.sum(1)`)."""
    assert proposals_idx.is_cuda and mask_scores.is_cuda and proposals_idx.dtype == torch.int32
    assert proposals_idx.is_contiguous() and mask_scores.stride(1) == 1 and mask_scores.dtype == torch.float32
    S = proposals_idx.size(0)
    npoint = torch.empty((n_proposals, n_classes), dtype=torch.int32, device=proposals_idx.device)
    with profiler.record('inst_count', 8 * S + 4 * S * n_classes):
        check(_lib.lib().sgb_inst_count(ptr(proposals_idx), ptr(mask_scores), mask_scores.stride(0), S, n_classes, float(thr),
                                        ptr(npoint), n_proposals, _stream()), 'sgb_inst_count')
    return npoint


def instance_bitmaps(proposals_idx, mask_scores, n_classes, thr, keep, n_points):
    """keep bool [nP, nI] -> (bitmaps int32 [n_inst, W], kept class int64 [n_inst], kept proposal int64 [n_inst]) in the
    reference's instance order (class-major, proposal order; softgroup.py:551-603)."""
    dev = proposals_idx.device
    nP = keep.size(0)
    kflat = keep.t().contiguous().view(-1)  # class-major
    slot = (torch.cumsum(kflat.int(), 0) - 1).int()
    slot = torch.where(kflat, slot, torch.full_like(slot, -1)).contiguous()
    kc, kp = keep.t().nonzero(as_tuple=True)
    n_inst = int(kc.numel())  # one small read-back: the bitmap allocation needs it
    W = bitmap_words(n_points)
    bm = torch.empty((n_inst, W), dtype=torch.int32, device=dev)
    with profiler.record('inst_scatter', 8 * proposals_idx.size(0) + 4 * n_inst * W):
        check(_lib.lib().sgb_inst_scatter(ptr(proposals_idx), ptr(mask_scores), mask_scores.stride(0), proposals_idx.size(0),
                                          n_classes, nP, float(thr), ptr(slot), ptr(bm), n_inst, int(n_points), _stream()),
              'sgb_inst_scatter')
    return bm, kc, kp


def bitmaps_from_pairs(rows, points, n_rows, n_points):
    """(row, point) int32 pairs -> bitmaps int32 [n_rows, W] (rows < 0 are skipped)."""
    assert rows.is_cuda and rows.dtype == torch.int32 and points.dtype == torch.int32
    W = bitmap_words(n_points)
    bm = torch.empty((n_rows, W), dtype=torch.int32, device=rows.device)
    check(_lib.lib().sgb_bitmap_set(ptr(rows.contiguous()), ptr(points.contiguous()), rows.numel(), ptr(bm), n_rows, int(n_points), 1,
                                    _stream()), 'sgb_bitmap_set')
    return bm


def bitmaps_to_rle(bitmaps, n_points):
    """bitmaps int32 [n_inst, W] (CUDA) -> list of the reference's RLE dicts (rle.py:5-19). Run boundaries are found on
    the GPU; only the transition positions (2 ints per run) cross PCIe; the strings are formatted by the host routine."""
    n_inst = bitmaps.size(0)
    if n_inst == 0:
        return []
    L = _lib.lib()
    dev = bitmaps.device
    ws = _ws(L.sgb_rle_workspace_bytes(n_inst, int(n_points)), dev)
    with profiler.record('rle_runs', 8 * bitmaps.numel()):
        total = check(L.sgb_rle_count(ptr(bitmaps), n_inst, int(n_points), ptr(ws), ws.numel(), _stream()), 'sgb_rle_count')
        trans = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
        inst_off = torch.empty(n_inst + 1, dtype=torch.int32, device=dev)
        check(L.sgb_rle_fill(ptr(bitmaps), n_inst, int(n_points), total, ptr(trans), ptr(inst_off), ptr(ws), ws.numel(), _stream()),
              'sgb_rle_fill')
    h_trans = torch.empty(max(total, 1), dtype=torch.int32, pin_memory=True)
    h_off = torch.empty(n_inst + 1, dtype=torch.int32, pin_memory=True)
    h_trans.copy_(trans, non_blocking=True)
    h_off.copy_(inst_off, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return format_runs(h_trans.numpy(), h_off.numpy(), n_points)


def format_runs(trans, offs, length):
    """Host: transition positions (1-based, ascending per mask) + offsets int32 [n+1] -> list of rle dicts."""
    trans = np.ascontiguousarray(trans, dtype=np.int32)
    offs = np.ascontiguousarray(offs, dtype=np.int32)
    n = offs.size - 1
    if n <= 0:
        return []
    cap = int(offs[-1]) * 12 + 64 * n + 64
    out = np.empty(cap, dtype=np.uint8)
    out_offs = np.empty(n + 1, dtype=np.int64)
    written = check(
        _lib.lib().sgb_rle_format_runs(trans.ctypes.data_as(ctypes.c_void_p), offs.ctypes.data_as(ctypes.c_void_p), n,
                                       out.ctypes.data_as(ctypes.c_void_p), cap, out_offs.ctypes.data_as(ctypes.c_void_p)),
        'sgb_rle_format_runs')
    buf = out[:written].tobytes()
    o = out_offs.tolist()
    return [dict(length=int(length), counts=buf[o[k]:o[k + 1]].decode('ascii')) for k in range(n)]


def bitmap_intersections(bitmaps, gslot, n_gt, n_points):
    """-> (inter int32 [n_rows, n_gt], vert int32 [n_rows], void int32 [n_rows]); gslot int32 [N]: gt column, -2 void, -1 other."""
    n_rows = bitmaps.size(0)
    dev = bitmaps.device
    inter = torch.empty((n_rows, max(n_gt, 1)), dtype=torch.int32, device=dev)
    vert = torch.empty(n_rows, dtype=torch.int32, device=dev)
    void = torch.empty(n_rows, dtype=torch.int32, device=dev)
    check(_lib.lib().sgb_bitmap_intersections(ptr(bitmaps), n_rows, int(n_points), ptr(gslot.contiguous()), int(n_gt), ptr(inter),
                                              ptr(vert), ptr(void), _stream()), 'sgb_bitmap_intersections')
    return inter[:, :n_gt], vert, void


def panoptic_paste(bitmaps, order, cls_values, semantic_preds, skip_iou, n_points):
    """softgroup.py:606-632 on bitmaps: -> (panoptic_cls uint32-valued int64 [N], panoptic_ids int64 [N]) CUDA tensors."""
    dev = bitmaps.device
    W = bitmap_words(n_points)
    pan_cls = semantic_preds.to(torch.int32).contiguous().clone()
    pan_ids = torch.zeros(int(n_points), dtype=torch.int32, device=dev)
    prev = torch.empty(W, dtype=torch.int32, device=dev)
    check(_lib.lib().sgb_panoptic_paste(ptr(bitmaps), int(n_points), ptr(order.int().contiguous()), ptr(cls_values.int().contiguous()),
                                        int(order.numel()), float(skip_iou), ptr(prev), ptr(pan_cls), ptr(pan_ids), _stream()),
          'sgb_panoptic_paste')
    return pan_cls, pan_ids
