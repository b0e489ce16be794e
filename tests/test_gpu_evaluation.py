"""N4: the GPU assignment step of the instance evaluation against the reference's own
`ScanNetEval.assign_instances_for_scan` / `evaluate_matches` (softgroup/evaluation/instance_eval.py:39-309, staged
verbatim by oracle/build_ref.py; the numpy >= 1.24 aliases are installed instead of editing it)."""
import numpy as np
import pytest

from softgroup_b200 import evaluation as sgb_eval
from softgroup_b200.util import rle_encode

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ref_eval():
    from oracle import build_ref
    m = build_ref.import_reference_eval()
    if m is None:
        pytest.skip('reference evaluation module not staged')
    sgb_eval.install_numpy_aliases()
    return m


def _scan(seed, n=60000, n_gt=25, n_pred=120, n_classes=18):
    rng = np.random.RandomState(seed)
    gts = np.zeros(n, np.int64)
    # ground truth: contiguous-ish blobs with class*1000 + instance id; some void classes (0 and 19+)
    pos = rng.permutation(n)
    cuts = np.sort(rng.choice(np.arange(1, n), n_gt, replace=False))
    segs = np.split(pos, cuts)
    for k, seg in enumerate(segs):
        cls = rng.randint(0, n_classes + 3)  # 0 and > n_classes are void
        gts[seg] = cls * 1000 + k + 1 if cls > 0 else 0
    preds = []
    for k in range(n_pred):
        seg = segs[rng.randint(len(segs))]
        keep = seg[rng.rand(seg.size) < rng.uniform(0.2, 1.0)]
        extra = rng.choice(n, rng.randint(0, 300), replace=False)
        m = np.zeros(n, np.uint8)
        m[keep] = 1
        m[extra] = 1
        if k % 17 == 0:
            m[:] = 0
            m[rng.choice(n, 30, replace=False)] = 1  # below min_region_size: skipped
        label = int(rng.randint(1, n_classes + 2))  # n_classes + 1 is not a valid label: skipped
        mask = rle_encode(m) if k % 2 else m  # both wire formats
        preds.append(dict(scan_id='scan%d' % seed, label_id=label, conf=float(rng.rand()), pred_mask=mask))
    return preds, gts


@pytest.mark.parametrize('use_label', [True, False])
def test_assign_instances_matches_reference(ref_eval, use_label):
    names = ['c%d' % i for i in range(18)]
    ev = ref_eval.ScanNetEval(names, use_label=use_label)
    preds, gts = _scan(1)
    want_g2p, want_p2g = ev.assign_instances_for_scan(preds, gts)
    got_g2p, got_p2g = sgb_eval.assign_instances_for_scan(ev, preds, gts)
    assert got_p2g.keys() == want_p2g.keys() and got_g2p.keys() == want_g2p.keys()
    n_match = 0
    for label in want_p2g:
        assert got_p2g[label] == want_p2g[label], label
        n_match += sum(len(p['matched_gt']) for p in want_p2g[label])
    for label in want_g2p:
        assert got_g2p[label] == want_g2p[label], label
    assert n_match > 30


def test_evaluate_matches_reference_averages(ref_eval):
    names = ['c%d' % i for i in range(18)]
    ev = ref_eval.ScanNetEval(names)
    scans = [_scan(s, n=30000, n_pred=80) for s in (2, 3)]
    matches = {}
    for i, (preds, gts) in enumerate(scans):
        g2p, p2g = ev.assign_instances_for_scan(preds, gts)
        matches['gt_%d' % i] = dict(gt=g2p, pred=p2g)
    ap, rc = ev.evaluate_matches(matches)
    want = ev.compute_averages(ap, rc)
    got = sgb_eval.evaluate(ev, [s[0] for s in scans], [s[1] for s in scans], verbose=False)
    assert got.keys() == want.keys()
    for k in want:
        if k == 'classes':
            assert got[k].keys() == want[k].keys()
            for c in want[k]:
                for kk in want[k][c]:
                    np.testing.assert_equal(got[k][c][kk], want[k][c][kk])
        else:
            np.testing.assert_equal(got[k], want[k])
    assert np.isfinite(want['all_ap_50%'])
