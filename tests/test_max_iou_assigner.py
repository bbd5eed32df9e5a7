"""Dense-anchor assignment: MaxIoUAssigner / bbox_overlaps.
CPU: the oracle restatement reproduces the fixture recorded from the real reference classes and the reference's own test vectors
(TOV_mmdetection/tests/test_utils/test_assigner.py:15-152).  GPU: ptb_max_iou_assign / ptb_bbox_overlaps, bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import anchors as oa

MAX_IOU_CFGS = oa.MAX_IOU_CFGS

BOXES = torch.FloatTensor([[0, 0, 10, 10], [10, 10, 20, 20], [5, 5, 15, 15], [32, 32, 38, 42]])
GTS = torch.FloatTensor([[0, 0, 10, 9], [0, 10, 10, 19]])


def test_oracle_matches_reference_golden_and_kats(golden_dir):
    gold = np.load(os.path.join(golden_dir, 'max_iou_assigner.npz'))
    for seed in (1, 2):
        a, g, l, ign = oa.synth_anchor_case(seed)
        assert np.array_equal(oa.bbox_overlaps(g, a).flatten()[::7].numpy(), gold[f's{seed}_iou_sub'])
        for ci, kw in enumerate(MAX_IOU_CFGS):
            gi, mo, lb = oa.max_iou_assign(a, g, l, ign, **kw)
            assert np.array_equal(gi.numpy(), gold[f's{seed}_c{ci}_gt_inds'].astype(np.int64)), (seed, ci)
            assert np.array_equal(lb.numpy(), gold[f's{seed}_c{ci}_labels'].astype(np.int64))
            assert np.array_equal(mo.numpy(), gold[f's{seed}_c{ci}_max_overlaps'])
    # the reference's own vectors
    gi, _, lb = oa.max_iou_assign(BOXES, GTS, torch.LongTensor([2, 3]), pos_iou_thr=0.5, neg_iou_thr=0.5)
    assert gi.tolist() == [1, 0, 2, 0] and len(lb) == 4
    boxes_i = BOXES.clone(); boxes_i[3] = torch.tensor([30., 32., 40., 42.])
    gi, _, _ = oa.max_iou_assign(boxes_i, GTS, None, torch.Tensor([[30, 30, 40, 40]]), pos_iou_thr=0.5, neg_iou_thr=0.5, ignore_iof_thr=0.5,
                                 ignore_wrt_candidates=False)
    assert gi.tolist() == [1, 0, 2, -1]
    gi, _, _ = oa.max_iou_assign(BOXES, torch.empty(0, 4), pos_iou_thr=0.5, neg_iou_thr=0.5)
    assert gi.tolist() == [0, 0, 0, 0]
    gi, _, lb = oa.max_iou_assign(torch.empty(0, 4), GTS, torch.LongTensor([2, 3]), pos_iou_thr=0.5, neg_iou_thr=0.5)
    assert len(gi) == 0 and tuple(lb.shape) == (0,)
    gi, _, lb = oa.max_iou_assign(torch.empty(0, 4), torch.empty(0, 4), pos_iou_thr=0.5, neg_iou_thr=0.5)
    assert len(gi) == 0 and lb is None


@pytest.fixture(scope='module')
def ops():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from pointtinybenchmark_b200 import ops
    return ops


@pytest.mark.gpu
def test_max_iou_assign_kernel_bit_exact(ops, golden_dir):
    from pointtinybenchmark_b200.assigners import MaxIoUAssigner
    dev = torch.device('cuda:0')
    gold = np.load(os.path.join(golden_dir, 'max_iou_assigner.npz'))
    for seed in (1, 2, 3):
        a, g, l, ign = oa.synth_anchor_case(seed)
        assert torch.equal(ops.bbox_overlaps(g.to(dev), a.to(dev)).cpu(), oa.bbox_overlaps(g, a)), 'IoU matrix'
        assert torch.equal(ops.bbox_overlaps(a.to(dev), ign.to(dev), 'iof').cpu(), oa.bbox_overlaps(a, ign, 'iof')), 'IoF matrix'
        for ci, kw in enumerate(MAX_IOU_CFGS):
            r = MaxIoUAssigner(**kw).assign(a.to(dev), g.to(dev), gt_bboxes_ignore=ign.to(dev), gt_labels=l.to(dev))
            gi, mo, lb = oa.max_iou_assign(a, g, l, ign, **kw)
            assert torch.equal(r.gt_inds.cpu(), gi), (seed, ci, int((r.gt_inds.cpu() != gi).sum()))
            assert torch.equal(r.labels.cpu(), lb) and torch.equal(r.max_overlaps.cpu(), mo)
            if seed < 3:
                assert np.array_equal(r.gt_inds.cpu().numpy(), gold[f's{seed}_c{ci}_gt_inds'].astype(np.int64)), 'vs golden'
    # the reference's own vectors and the empty cases
    A = MaxIoUAssigner(pos_iou_thr=0.5, neg_iou_thr=0.5)
    r = A.assign(BOXES.to(dev), GTS.to(dev), gt_labels=torch.LongTensor([2, 3]).to(dev))
    assert r.gt_inds.tolist() == [1, 0, 2, 0] and len(r.labels) == 4 and r.num_gts == 2
    boxes_i = BOXES.clone(); boxes_i[3] = torch.tensor([30., 32., 40., 42.])
    r = MaxIoUAssigner(0.5, 0.5, ignore_iof_thr=0.5, ignore_wrt_candidates=False).assign(
        boxes_i.to(dev), GTS.to(dev), gt_bboxes_ignore=torch.Tensor([[30, 30, 40, 40]]).to(dev))
    assert r.gt_inds.tolist() == [1, 0, 2, -1]
    assert A.assign(BOXES.to(dev), torch.empty(0, 4, device=dev)).gt_inds.tolist() == [0, 0, 0, 0]
    r = A.assign(torch.empty(0, 4, device=dev), GTS.to(dev), gt_labels=torch.LongTensor([2, 3]).to(dev))
    assert len(r.gt_inds) == 0 and tuple(r.labels.shape) == (0,)
    assert len(A.assign(torch.empty(0, 4, device=dev), torch.empty(0, 4, device=dev)).gt_inds) == 0
    with pytest.raises(RuntimeError, match='no CPU'):
        A.assign(BOXES, GTS)


@pytest.mark.gpu
def test_max_iou_assign_config4_size(ops):
    """BASELINE.json configs[3] size: 81 840 anchors (640x512 tile, 5 levels, 3 anchors per cell) x 300 GTs (more than one GT tile):
    against the oracle, deterministic, and consistent with the explicit IoU matrix."""
    dev = torch.device('cuda:0')
    a, g, l, ign = oa.synth_anchor_case(11, n_anchor=81840, n_gt=300, n_ign=5)
    kw = dict(pos_iou_thr=0.7, neg_iou_thr=0.3, min_pos_iou=0.3, match_low_quality=True, ignore_iof_thr=0.5)
    gi, mo, lb = ops.max_iou_assign(a.to(dev), g.to(dev), l.to(dev), ign.to(dev), **kw)
    gi2, _, _ = ops.max_iou_assign(a.to(dev), g.to(dev), l.to(dev), ign.to(dev), **kw)
    assert torch.equal(gi, gi2)
    ogi, omo, olb = oa.max_iou_assign(a, g, l, ign, **kw)
    assert torch.equal(gi.cpu(), ogi) and torch.equal(mo.cpu(), omo) and torch.equal(lb.cpu(), olb)
    M = ops.bbox_overlaps(g.to(dev), a.to(dev))
    keep = mo >= 0
    assert torch.equal(M.max(dim=0)[0][keep], mo[keep])
    print(f'[config 4 size] positives {int((gi > 0).sum())}, negatives {int((gi == 0).sum())}, ignored {int((gi < 0).sum())}')
