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


# ---------------------------------------------------------------------------------------------------------------------------
# RPN proposal path: AnchorGenerator grid anchors, DeltaXYWHBBoxCoder.decode, per-level top-k, batched NMS.
# Pinned: oracle/make_golden.py::golden_rpn runs the real reference RPNHead.get_bboxes (through the mmcv stub) on seeded inputs,
# asserts equality with this restatement and stores tests/golden/rpn_proposals.npz.
# ---------------------------------------------------------------------------------------------------------------------------
import numpy as np


def base_anchors(base_size, scales, ratios, center_offset=0.0, scale_major=True):
    """ref anchor_generator.py:142-185 (gen_single_level_base_anchors): (A,4) fp32, A = len(ratios)*len(scales)."""
    scales, ratios = torch.as_tensor(scales, dtype=torch.float32), torch.as_tensor(ratios, dtype=torch.float32)
    w = h = base_size
    xc, yc = center_offset * w, center_offset * h
    hr = torch.sqrt(ratios)
    wr = 1 / hr
    if scale_major:
        ws = (w * wr[:, None] * scales[None, :]).view(-1)
        hs = (h * hr[:, None] * scales[None, :]).view(-1)
    else:
        ws = (w * scales[:, None] * wr[None, :]).view(-1)
        hs = (h * scales[:, None] * hr[None, :]).view(-1)
    return torch.stack([xc - 0.5 * ws, yc - 0.5 * hs, xc + 0.5 * ws, yc + 0.5 * hs], dim=-1)


def grid_anchors(base, feat_hw, stride_wh):
    """ref anchor_generator.py:233-270: (H*W*A, 4), cell-major then anchor (index q = (y*W + x)*A + a)."""
    fh, fw = feat_hw
    sx = torch.arange(0, fw) * stride_wh[0]
    sy = torch.arange(0, fh) * stride_wh[1]
    xx = sx.repeat(fh)
    yy = sy.view(-1, 1).repeat(1, fw).view(-1)
    shifts = torch.stack([xx, yy, xx, yy], dim=-1).type_as(base)
    return (base[None, :, :] + shifts[:, None, :]).view(-1, 4)


def delta2bbox(rois, deltas, means=(0., 0., 0., 0.), stds=(1., 1., 1., 1.), max_shape=None, wh_ratio_clip=16 / 1000):
    """ref delta_xywh_bbox_coder.py:144-270 (clip_border=True, add_ctr_clamp=False); rois (..., 4), deltas (..., 4),
    max_shape (h, w) or a (B, 2) tensor for batched rois."""
    means = deltas.new_tensor(means).view(1, -1)
    stds = deltas.new_tensor(stds).view(1, -1)
    d = deltas * stds + means
    dx, dy, dw, dh = d[..., 0::4], d[..., 1::4], d[..., 2::4], d[..., 3::4]
    x1, y1, x2, y2 = rois[..., 0], rois[..., 1], rois[..., 2], rois[..., 3]
    px = ((x1 + x2) * 0.5).unsqueeze(-1).expand_as(dx)
    py = ((y1 + y2) * 0.5).unsqueeze(-1).expand_as(dy)
    pw = (x2 - x1).unsqueeze(-1).expand_as(dw)
    ph = (y2 - y1).unsqueeze(-1).expand_as(dh)
    dxw, dyh = pw * dx, ph * dy
    max_ratio = np.abs(np.log(wh_ratio_clip))
    dw = dw.clamp(min=-max_ratio, max=max_ratio)
    dh = dh.clamp(min=-max_ratio, max=max_ratio)
    gw, gh = pw * dw.exp(), ph * dh.exp()
    gx, gy = px + dxw, py + dyh
    b = torch.stack([gx - gw * 0.5, gy - gh * 0.5, gx + gw * 0.5, gy + gh * 0.5], dim=-1).view(deltas.size())
    if max_shape is not None:
        ms = max_shape if isinstance(max_shape, torch.Tensor) else b.new_tensor(max_shape)
        ms = ms[..., :2].type_as(b)
        mn = b.new_tensor(0)
        mx = torch.cat([ms] * (deltas.size(-1) // 2), dim=-1).flip(-1).unsqueeze(-2)
        b = torch.where(b < mn, mn, b)
        b = torch.where(b > mx, mx, b)
    return b


RPN_CFG = dict(scales=[2], ratios=[0.5, 1.0, 2.0], strides=[4, 8, 16, 32, 64], means=(0., 0., 0., 0.), stds=(1., 1., 1., 1.),
               nms_pre=1000, max_per_img=1000, iou_threshold=0.7, min_bbox_size=0)     # faster_rcnn_r50_fpn_1x_TinyPerson640.py:25-40,106-112


def rpn_proposals(cls_scores, bbox_preds, img_shapes, cfg=RPN_CFG, return_all=False):
    """ref rpn_head.py:78-186 (RPNHead._get_bboxes, use_sigmoid_cls=True) after anchor_head.py:551-590 (grid anchors per level).
    cls_scores[l] (B, A, H_l, W_l), bbox_preds[l] (B, 4A, H_l, W_l), img_shapes [(h, w, 3)] -> per image dets (m, 5)."""
    from oracle import p2p as op2p
    B = cls_scores[0].shape[0]
    lvl_scores, lvl_preds, lvl_anchors, lvl_ids, lvl_idx = [], [], [], [], []
    for l, (cs, bp) in enumerate(zip(cls_scores, bbox_preds)):
        H, W = cs.shape[-2:]
        s = cfg['strides'][l]
        anchors = grid_anchors(base_anchors(s, cfg['scales'], cfg['ratios']), (H, W), (s, s))
        scores = cs.permute(0, 2, 3, 1).reshape(B, -1).sigmoid()
        pred = bp.permute(0, 2, 3, 1).reshape(B, -1, 4)
        anchors = anchors.expand_as(pred)
        idx = torch.arange(scores.shape[1])[None].expand(B, -1)
        if cfg['nms_pre'] > 0 and pred.size(1) > cfg['nms_pre']:
            ranked, rank_inds = scores.sort(descending=True)
            idx = rank_inds[:, :cfg['nms_pre']]
            scores = ranked[:, :cfg['nms_pre']]
            bi = torch.arange(B).view(-1, 1).expand_as(idx)
            pred = pred[bi, idx, :]
            anchors = anchors[bi, idx, :]
        lvl_scores.append(scores); lvl_preds.append(pred); lvl_anchors.append(anchors); lvl_idx.append(idx)
        lvl_ids.append(scores.new_full((B, scores.size(1)), l, dtype=torch.long))
    scores = torch.cat(lvl_scores, 1)
    props = delta2bbox(torch.cat(lvl_anchors, 1), torch.cat(lvl_preds, 1), cfg['means'], cfg['stds'], max_shape=img_shapes)
    ids = torch.cat(lvl_ids, 1)
    out, allv = [], []
    for b in range(B):
        p, sc, li = props[b], scores[b], ids[b]
        pos = torch.arange(p.shape[0])
        if cfg['min_bbox_size'] >= 0:
            w, h = p[:, 2] - p[:, 0], p[:, 3] - p[:, 1]
            v = torch.nonzero((w > cfg['min_bbox_size']) & (h > cfg['min_bbox_size']), as_tuple=False).squeeze(1)
            if v.sum().item() != len(p):                        # (sic) rpn_head.py:178
                p, sc, li, pos = p[v, :], sc[v], li[v], pos[v]
        dets, keep = op2p.batched_nms(p, sc, li, cfg['iou_threshold'])
        out.append(dets[:cfg['max_per_img']])
        allv.append(dict(keep_pos=pos[keep][:cfg['max_per_img']], levels=li[keep][:cfg['max_per_img']]))
    if return_all:
        return out, dict(cand_idx=torch.cat(lvl_idx, 1), cand_boxes=props, cand_scores=scores, per_image=allv)
    return out


def synth_rpn_inputs(seed, B=2, size=(512, 640), A=3, strides=(4, 8, 16, 32, 64), score_mu=-3.0, score_sigma=1.5):
    """seeded RPN outputs for a (h, w) = size tile: logits ~ N(mu, sigma) made pairwise distinct per level, deltas ~ N(0, 0.3)."""
    g = torch.Generator().manual_seed(seed)
    cls, box = [], []
    for s in strides:
        H, W = -(-size[0] // s), -(-size[1] // s)
        c = torch.randn(B, A, H, W, generator=g) * score_sigma + score_mu
        cls.append(c)
        box.append(torch.randn(B, 4 * A, H, W, generator=g) * 0.3)
    img_shapes = [(size[0] - 3 * b, size[1] - 5 * b, 3) for b in range(B)]
    return cls, box, img_shapes
