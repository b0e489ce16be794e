"""Generate tests/golden/*.npz from the UNMODIFIED compiled reference (oracle/_ref, built by
oracle/build_ref.py from /root/reference/softgroup/ops/src). Run in the build container only:

    python tests/golden/make_golden.py

The reference has CPU implementations of voxelize_idx, bfs_cluster and build_and_export_octree only
(SURVEY.md 8c); neighbour lists fed to bfs_cluster come from the oracle's brute-force ball query
(and from random graphs), since the fixture pins bfs_cluster itself, not the lists.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle.build_ref import build, load_ref  # noqa: E402
from softgroup_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def ref_voxelize(ref, coords, batch, mode):
    c = torch.from_numpy(coords)
    oc = c.new()
    im = torch.IntTensor(c.size(0)).zero_()
    om = im.new()
    ref.voxelize_idx(c, oc, im, om, batch, mode)
    return oc.numpy(), im.numpy(), om.numpy()


def ref_bfs(ref, mean, idx, sl, thr, cls):
    ci, co = torch.IntTensor(), torch.IntTensor()
    ref.bfs_cluster(torch.from_numpy(mean), torch.from_numpy(idx), torch.from_numpy(sl), ci, co, sl.shape[0],
                    float(thr), int(cls))
    return ci.numpy().reshape(-1, 2), co.numpy()


def main():
    build()
    ref = load_ref()
    assert ref is not None, 'reference extension unavailable'
    g = {}
    # ---- voxelize_idx: C1 scan (2k points), mode 4; ragged 2-batch case, modes 1,2,3 -------------------------
    scan = synth.make_scan('c1_plumbing', seed=0)
    g['vox_c1_coords'] = scan['coords']
    oc, im, om = ref_voxelize(ref, scan['coords'], 1, 4)
    g['vox_c1_out_coords'], g['vox_c1_input_map'], g['vox_c1_output_map'] = oc, im, om
    rng = np.random.RandomState(7)
    rag = np.concatenate([np.zeros((300, 1), np.int64), rng.randint(0, 6, (300, 3))], 1)
    rag2 = np.concatenate([np.ones((57, 1), np.int64), rng.randint(-3, 4, (57, 3))], 1)  # negative coords (pyramid_map)
    rag = np.concatenate([rag, rag2], 0)
    g['vox_rag_coords'] = rag
    for mode in (1, 2, 3, 4):
        oc, im, om = ref_voxelize(ref, rag, 2, mode)
        g['vox_rag_m%d_out_coords' % mode], g['vox_rag_m%d_input_map' % mode], g['vox_rag_m%d_output_map' % mode] = oc, im, om
    # ---- bfs_cluster: per-class lists of the C1 scan -----------------------------------------------------------
    scores, off = synth.grouping_inputs(scan, sigma=0.03, seed=0)
    sem = scan['semantic_labels']
    mean = np.array([-1., -1., 3917., 12056., 2303., 8331., 3948., 3166., 5629., 11719., 1003., 3317., 4912., 10221.,
                     3889., 4136., 2120., 945., 3967., 2589.], np.float32)  # configs/softgroup/softgroup_scannet.yaml:13-17
    k = 0
    for cls in np.unique(sem):
        if cls < 2:
            continue
        sel = np.where(sem == cls)[0]
        xyz = (scan['coords_float'][sel] + off[sel]).astype(np.float32)
        bi = np.zeros(len(sel), np.int32)
        bo = np.array([0, len(sel)], np.int32)
        idx, sl = oracle.ballquery_batch_p(xyz, bi, bo, 0.04)
        for thr in (0.05, 0.01):
            ci, co = ref_bfs(ref, mean, idx, sl, thr, cls)
            g['bfs%d_xyz' % k], g['bfs%d_idx' % k], g['bfs%d_sl' % k] = xyz, idx, sl
            g['bfs%d_thr' % k], g['bfs%d_cls' % k] = np.float32(thr), np.int32(cls)
            g['bfs%d_cidx' % k], g['bfs%d_coff' % k] = ci, co
            k += 1
    # random directed graphs: asymmetric + permuted list order (octree-like) + absolute threshold (mean == -1)
    for t in range(6):
        n = 40 + 30 * t
        lens = rng.randint(0, 7, n)
        start = np.concatenate([[0], np.cumsum(lens)[:-1]])
        idx = np.concatenate([rng.choice(n, l, replace=False) for l in lens] + [np.zeros(0, np.int64)]).astype(np.int32)
        sl = np.stack([start, lens], 1).astype(np.int32)
        if idx.size == 0:
            idx = np.zeros(1, np.int32)
        ci, co = ref_bfs(ref, mean, idx, sl, 2 + t, 0)
        g['bfs%d_xyz' % k] = np.zeros((n, 3), np.float32)
        g['bfs%d_idx' % k], g['bfs%d_sl' % k] = idx, sl
        g['bfs%d_thr' % k], g['bfs%d_cls' % k] = np.float32(2 + t), np.int32(0)
        g['bfs%d_cidx' % k], g['bfs%d_coff' % k] = ci, co
        k += 1
    g['bfs_count'] = np.int32(k)
    g['class_numpoint_mean'] = mean
    # ---- octree build ------------------------------------------------------------------------------------------
    pts = (rng.rand(700, 3) * np.array([3., 2., 1.])).astype(np.float32)
    mx, mn = pts.max(0), pts.min(0)
    xyzwhl = torch.from_numpy(np.concatenate([(mx + mn) / np.float32(2), mx - mn]).astype(np.float32))
    boxes = torch.zeros((585, 6), dtype=torch.float32)
    pt_inds = torch.zeros(700, dtype=torch.int32)
    psl = torch.zeros((512, 2), dtype=torch.int32)
    ref.build_and_export_octree(torch.from_numpy(pts), xyzwhl, boxes, pt_inds, psl, 3)
    g['oct_pts'], g['oct_boxes'], g['oct_pt_inds'], g['oct_psl'] = pts, boxes.numpy(), pt_inds.numpy(), psl.numpy()
    np.savez_compressed(os.path.join(OUT, 'ref_ops_golden.npz'), **g)
    print('wrote', os.path.join(OUT, 'ref_ops_golden.npz'), 'bfs cases', k)


if __name__ == '__main__':
    main()
