"""MaxIoUAssigner — host-side mirror of mmdet/core/bbox/assigners/max_iou_assigner.py:9-212 over ptb_max_iou_assign
(the dense-anchor assignment of BASELINE.json configs[3]).  Same ctor kwargs, `assign` signature and result
fields (`num_gts`, `gt_inds`, `max_overlaps`, `labels`) as the reference's AssignResult (assign_result.py:42-46)."""
from . import ops


class AssignResult:
    def __init__(self, num_gts, gt_inds, max_overlaps, labels=None):
        self.num_gts, self.gt_inds, self.max_overlaps, self.labels = num_gts, gt_inds, max_overlaps, labels

    @property
    def num_preds(self):
        return len(self.gt_inds)


class MaxIoUAssigner:
    def __init__(self, pos_iou_thr, neg_iou_thr, min_pos_iou=.0, gt_max_assign_all=True, ignore_iof_thr=-1, ignore_wrt_candidates=True,
                 match_low_quality=True, gpu_assign_thr=-1, iou_calculator=None):
        if iou_calculator is not None and dict(iou_calculator).get('type', 'BboxOverlaps2D') != 'BboxOverlaps2D':
            raise NotImplementedError(f'iou_calculator {iou_calculator}')
        if iou_calculator is not None and dict(iou_calculator).get('dtype') == 'fp16':
            raise NotImplementedError('BboxOverlaps2D(dtype=fp16)')
        self.pos_iou_thr, self.neg_iou_thr, self.min_pos_iou = pos_iou_thr, neg_iou_thr, min_pos_iou
        self.gt_max_assign_all, self.ignore_iof_thr, self.ignore_wrt_candidates = gt_max_assign_all, ignore_iof_thr, ignore_wrt_candidates
        self.match_low_quality = match_low_quality
        self.gpu_assign_thr = gpu_assign_thr          # accepted for config compatibility: there is no CPU assignment path here

    def assign(self, bboxes, gt_bboxes, gt_bboxes_ignore=None, gt_labels=None):
        if not bboxes.is_cuda:
            raise RuntimeError('MaxIoUAssigner (B200) runs on CUDA tensors only; there is no CPU fallback')
        b = bboxes[:, :4].float().contiguous()
        g = gt_bboxes[:, :4].float().contiguous()
        ign = gt_bboxes_ignore[:, :4].float().contiguous() if gt_bboxes_ignore is not None else None
        gt_inds, max_ov, labels = ops.max_iou_assign(b, g, gt_labels, ign, self.pos_iou_thr, self.neg_iou_thr, self.min_pos_iou,
                                                     self.gt_max_assign_all, self.ignore_iof_thr, self.ignore_wrt_candidates,
                                                     self.match_low_quality)
        if g.shape[0] == 0 and gt_labels is None:
            labels = None
        return AssignResult(g.shape[0], gt_inds, max_ov, labels)
