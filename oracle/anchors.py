"""CPU restatement of the dense-anchor assignment path (BASELINE.json configs[3], Faster-RCNN style RPN):
`bbox_overlaps` (mmdet/core/bbox/iou_calculators/iou2d_calculator.py:76-262, modes iou / iof) and `MaxIoUAssigner`
(mmdet/core/bbox/assigners/max_iou_assigner.py:9-212).  Test infrastructure — only tests/, smoke() and bench's CPU legs may import
this.  Pinned: oracle/make_golden.py::golden_max_iou runs the real reference classes on seeded inputs, asserts equality and stores
tests/golden/max_iou_assigner.npz; the reference's own vectors (tests/test_utils/test_assigner.py:15-152) are in the tests."""
import torch


def bbox_overlaps(b1, b2, mode='iou', eps=1e-6):
    """ref iou2d_calculator.py:211-256 (is_aligned=False): (m,4),(n,4) -> (m,n); same torch ops in the same order."""
    rows, cols = b1.size(0), b2.size(0)
    if rows * cols == 0:
        return b1.new_zeros((rows, cols))
    area1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    area2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    lt = torch.max(b1[:, None, :2], b2[None, :, :2])
    rb = torch.min(b1[:, None, 2:], b2[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    overlap = wh[..., 0] * wh[..., 1]
    union = area1[:, None] + area2[None, :] - overlap if mode == 'iou' else area1[:, None].expand(rows, cols)
    union = torch.max(union, union.new_tensor([eps]))
    return overlap / union


def max_iou_assign(bboxes, gt_bboxes, gt_labels=None, gt_bboxes_ignore=None, pos_iou_thr=0.5, neg_iou_thr=0.5, min_pos_iou=0.0,
                   gt_max_assign_all=True, ignore_iof_thr=-1, ignore_wrt_candidates=True, match_low_quality=True):
    """ref max_iou_assigner.py:60-212.  returns gt_inds (N,) int64 (-1 ignore, 0 negative, i+1 positive), max_overlaps (N,),
    labels (N,) int64 or None."""
    bboxes, gt_bboxes = bboxes[:, :4], gt_bboxes[:, :4]
    overlaps = bbox_overlaps(gt_bboxes, bboxes)                                            # (k, n)
    if ignore_iof_thr > 0 and gt_bboxes_ignore is not None and gt_bboxes_ignore.numel() > 0 and bboxes.numel() > 0:
        if ignore_wrt_candidates:
            ign = bbox_overlaps(bboxes, gt_bboxes_ignore, mode='iof').max(dim=1)[0]
        else:
            ign = bbox_overlaps(gt_bboxes_ignore, bboxes, mode='iof').max(dim=0)[0]
        overlaps[:, ign > ignore_iof_thr] = -1
    k, n = overlaps.shape
    gt_inds = overlaps.new_full((n,), -1, dtype=torch.long)
    if k == 0 or n == 0:
        if k == 0:
            gt_inds[:] = 0
        return gt_inds, overlaps.new_zeros((n,)), (None if gt_labels is None else overlaps.new_full((n,), -1, dtype=torch.long))
    max_overlaps, argmax_overlaps = overlaps.max(dim=0)
    gt_max_overlaps, gt_argmax_overlaps = overlaps.max(dim=1)
    if isinstance(neg_iou_thr, float):
        gt_inds[(max_overlaps >= 0) & (max_overlaps < neg_iou_thr)] = 0
    elif isinstance(neg_iou_thr, tuple):
        gt_inds[(max_overlaps >= neg_iou_thr[0]) & (max_overlaps < neg_iou_thr[1])] = 0
    pos = max_overlaps >= pos_iou_thr
    gt_inds[pos] = argmax_overlaps[pos] + 1
    if match_low_quality:
        for i in range(k):
            if gt_max_overlaps[i] >= min_pos_iou:
                if gt_max_assign_all:
                    gt_inds[overlaps[i, :] == gt_max_overlaps[i]] = i + 1
                else:
                    gt_inds[gt_argmax_overlaps[i]] = i + 1
    labels = None
    if gt_labels is not None:
        labels = gt_inds.new_full((n,), -1)
        p = gt_inds > 0
        labels[p] = gt_labels[gt_inds[p] - 1]
    return gt_inds, max_overlaps, labels


# assigner configurations of the fixtures / parity tests (RPN-like, R-CNN-like, tuple thresholds + single-anchor matching, ignore regions)
MAX_IOU_CFGS = [dict(pos_iou_thr=0.7, neg_iou_thr=0.3, min_pos_iou=0.3, match_low_quality=True, ignore_iof_thr=-1),
                dict(pos_iou_thr=0.5, neg_iou_thr=0.5, min_pos_iou=0.5, match_low_quality=False, ignore_iof_thr=-1),
                dict(pos_iou_thr=0.7, neg_iou_thr=(0.1, 0.3), min_pos_iou=0.0, gt_max_assign_all=False, ignore_iof_thr=0.5),
                dict(pos_iou_thr=0.5, neg_iou_thr=0.4, min_pos_iou=0.0, ignore_iof_thr=0.3, ignore_wrt_candidates=False)]


def synth_anchor_case(seed, n_anchor=3000, n_gt=17, n_ign=3, size=(512, 640)):
    """seeded dense-anchor-like boxes + GT boxes; a few anchors coincide with GTs / share the maximum so the low-quality matching and
    the gt_max_assign_all tie rule are exercised."""
    g = torch.Generator().manual_seed(seed)
    h, w = size
    c = torch.rand(n_anchor, 2, generator=g) * torch.tensor([w, h], dtype=torch.float32)
    s = torch.tensor([8., 16., 32., 64.])[torch.randint(0, 4, (n_anchor,), generator=g)]
    r = torch.tensor([0.5, 1.0, 2.0])[torch.randint(0, 3, (n_anchor,), generator=g)]
    ws, hs = s * r.sqrt(), s / r.sqrt()
    anchors = torch.stack([c[:, 0] - ws / 2, c[:, 1] - hs / 2, c[:, 0] + ws / 2, c[:, 1] + hs / 2], 1)
    gc = torch.rand(n_gt, 2, generator=g) * torch.tensor([w, h], dtype=torch.float32)
    gs = torch.rand(n_gt, 2, generator=g) * 60 + 4
    gts = torch.cat([gc - gs / 2, gc + gs / 2], 1)
    if n_gt >= 2 and n_anchor >= 8:
        anchors[0] = gts[0]                                    # IoU 1 with gt 0
        anchors[1] = anchors[2] = gts[1] + torch.tensor([1., 0., 1., 0.])   # two anchors tie for gt 1's maximum
        gts[-1] = torch.tensor([w + 50., h + 50., w + 60., h + 60.])        # a GT no anchor touches (max overlap 0 >= min_pos_iou 0)
    ic = torch.rand(n_ign, 2, generator=g) * torch.tensor([w, h], dtype=torch.float32)
    ign = torch.cat([ic - 40, ic + 40], 1)
    labels = torch.randint(0, 5, (n_gt,), generator=g)
    return anchors, gts, labels, ign
