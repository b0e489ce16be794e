"""Run-length encoding of binary point masks in the reference's wire format
(softgroup/util/rle.py:5-19: dict(length=N, counts='start len start len ...'), 1-based starts)."""
import numpy as np


def rle_encode(mask):
    length = mask.shape[0]
    m = np.concatenate([[0], np.asarray(mask), [0]])
    runs = np.where(m[1:] != m[:-1])[0] + 1
    runs[1::2] -= runs[::2]
    return dict(length=length, counts=' '.join(map(str, runs.tolist())))


def rle_encode_ids(ids_sorted, length):
    """Same encoding from the ascending list of set positions (no dense mask)."""
    ids = np.asarray(ids_sorted, dtype=np.int64)
    if ids.size == 0:
        return dict(length=length, counts='')
    brk = np.where(np.diff(ids) != 1)[0]
    starts = np.concatenate([[ids[0]], ids[brk + 1]]) + 1
    ends = np.concatenate([ids[brk], [ids[-1]]]) + 1
    runs = np.empty(starts.size * 2, np.int64)
    runs[0::2] = starts
    runs[1::2] = ends - starts + 1
    return dict(length=length, counts=' '.join(map(str, runs.tolist())))


def rle_decode(rle):
    length = rle['length']
    s = rle['counts'].split()
    starts = np.asarray(s[0::2], dtype=np.int64) - 1
    nums = np.asarray(s[1::2], dtype=np.int64)
    mask = np.zeros(length, dtype=np.uint8)
    for lo, n in zip(starts, nums):
        mask[lo:lo + n] = 1
    return mask


def rle_encode_many(ids, offs, length):
    """Batch version through the library's host formatter: ids int32 (ascending inside each mask, masks back to
    back), offs int64 [n+1] -> list of rle dicts."""
    import ctypes
    from ..ops import _lib
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    offs = np.ascontiguousarray(offs, dtype=np.int64)
    n = offs.size - 1
    if n <= 0:
        return []
    cap = int(ids.size) * 24 + 64 * n + 64
    out = np.empty(cap, dtype=np.uint8)
    out_offs = np.empty(n + 1, dtype=np.int64)
    written = _lib.check(
        _lib.lib().sgb_rle_format_ids(ids.ctypes.data_as(ctypes.c_void_p), offs.ctypes.data_as(ctypes.c_void_p), n,
                                      out.ctypes.data_as(ctypes.c_void_p), cap,
                                      out_offs.ctypes.data_as(ctypes.c_void_p)), 'sgb_rle_format_ids')
    buf = out[:written].tobytes()
    o = out_offs.tolist()
    return [dict(length=length, counts=buf[o[k]:o[k + 1]].decode('ascii')) for k in range(n)]
