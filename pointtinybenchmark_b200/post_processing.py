"""multiclass_nms — host-side mirror of mmdet/core/post_processing/bbox_nms.py:7-94 (which calls the third-party
mmcv.ops.nms.batched_nms) over ptb_multiclass_nms_boxes / ptb_multiclass_soft_nms: same arguments and return values
(`dets (k,5)`, `labels (k,)`, optionally `keep` = indices into the score-filtered candidate list, as the reference returns them).
CUDA tensors only; limits of the kernel: n <= 4096 boxes, 0 < max_num <= 1024, shared boxes (n, 4)."""
import torch

from . import ops


def multiclass_nms(multi_bboxes, multi_scores, score_thr, nms_cfg, max_num=-1, score_factors=None, return_inds=False):
    if not multi_bboxes.is_cuda:
        raise RuntimeError('multiclass_nms (B200) runs on CUDA tensors only; there is no CPU fallback')
    if multi_bboxes.shape[1] != 4:
        raise NotImplementedError('class-specific boxes (n, #class*4)')
    if score_factors is not None:
        raise NotImplementedError('score_factors')
    if not (0 < max_num <= 1024):
        raise NotImplementedError('max_num must be in [1, 1024] (the reference default -1 = unlimited is not supported)')
    cfg = dict(nms_cfg)
    kind = cfg.pop('type', 'nms')
    if cfg.pop('class_agnostic', False):
        raise NotImplementedError('class_agnostic NMS')
    iou = cfg.pop('iou_threshold', cfg.pop('iou_thr', 0.5))
    boxes = multi_bboxes.float().contiguous()[None]
    scores = multi_scores[:, :-1].float().contiguous()[None]              # the last column is the background class
    if kind == 'nms':
        cnt, det, lab, keep, _ = ops.multiclass_nms_boxes(boxes, scores, score_thr, iou, max_num)
    elif kind == 'soft_nms':
        cnt, det, lab, keep, _ = ops.multiclass_soft_nms(boxes, scores, None, score_thr, iou, max_num, sigma=cfg.get('sigma', 0.5),
                                                         min_score=cfg.get('min_score', 1e-3), method=cfg.get('method', 'linear'))
    else:
        raise NotImplementedError(f'nms type {kind}')
    k = int(cnt[0])
    dets, labels = det[0, :k], lab[0, :k].long()
    return (dets, labels, keep[0, :k].long()) if return_inds else (dets, labels)
