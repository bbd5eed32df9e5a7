"""Dense-anchor proposal path (BASELINE.json configs[3]) — host-side mirror of the reference pieces around
ptb_rpn_proposals:

  AnchorGenerator   mmdet/core/anchor/anchor_generator.py:9-330 (base anchors, grid anchors, valid flags; same ctor kwargs).  The base
                    anchors are input-independent (a few floats per level): computed once with torch CPU ops in the reference's
                    order, so they are bit-identical; the H*W*A grid anchors are NOT materialised on the proposal path — the decode
                    kernel forms `base[a] + shift(x, y)` on the fly — `grid_anchors` exists for callers that want the tensor
                    (e.g. to feed MaxIoUAssigner).
  RPNProposals      the `get_bboxes` of RPNHead (AnchorHead.get_bboxes, anchor_head.py:551-590 -> RPNHead._get_bboxes,
                    rpn_head.py:78-186): same arguments (cls_scores, bbox_preds, img_metas, cfg, rescale, with_nms), same result
                    (list of (n, 5) tensors), one library call for the whole batch instead of the per-level / per-image Python loop.
There is no CPU path: CUDA tensors only.
"""
import numpy as np
import torch

from . import ops
from .registry import CfgNode


def _pair(v):
    return (v, v) if not isinstance(v, (tuple, list)) else tuple(v)


class AnchorGenerator:
    def __init__(self, strides, ratios, scales=None, base_sizes=None, scale_major=True, octave_base_scale=None, scales_per_octave=None,
                 centers=None, center_offset=0.):
        if center_offset != 0:
            assert centers is None, f'center cannot be set when center_offset != 0, {centers} is given.'
        if not (0 <= center_offset <= 1):
            raise ValueError(f'center_offset should be in range [0, 1], {center_offset} is given.')
        if centers is not None:
            assert len(centers) == len(strides)
        self.strides = [_pair(s) for s in strides]
        self.base_sizes = [min(s) for s in self.strides] if base_sizes is None else list(base_sizes)
        assert len(self.base_sizes) == len(self.strides)
        assert (octave_base_scale is not None and scales_per_octave is not None) ^ (scales is not None), \
            'scales and octave_base_scale with scales_per_octave cannot be set at the same time'
        if scales is not None:
            self.scales = torch.Tensor(scales)
        else:
            octave_scales = np.array([2 ** (i / scales_per_octave) for i in range(scales_per_octave)])
            self.scales = torch.Tensor(octave_scales * octave_base_scale)
        self.octave_base_scale, self.scales_per_octave = octave_base_scale, scales_per_octave
        self.ratios = torch.Tensor(ratios)
        self.scale_major, self.centers, self.center_offset = scale_major, centers, center_offset
        self.base_anchors = self.gen_base_anchors()

    @property
    def num_base_anchors(self):
        return [b.size(0) for b in self.base_anchors]

    @property
    def num_levels(self):
        return len(self.strides)

    def gen_base_anchors(self):
        return [self.gen_single_level_base_anchors(bs, self.scales, self.ratios, None if self.centers is None else self.centers[i])
                for i, bs in enumerate(self.base_sizes)]

    def gen_single_level_base_anchors(self, base_size, scales, ratios, center=None):
        w = h = base_size
        xc, yc = (self.center_offset * w, self.center_offset * h) if center is None else center
        hr = torch.sqrt(ratios)
        wr = 1 / hr
        if self.scale_major:
            ws, hs = (w * wr[:, None] * scales[None, :]).view(-1), (h * hr[:, None] * scales[None, :]).view(-1)
        else:
            ws, hs = (w * scales[:, None] * wr[None, :]).view(-1), (h * scales[:, None] * hr[None, :]).view(-1)
        return torch.stack([xc - 0.5 * ws, yc - 0.5 * hs, xc + 0.5 * ws, yc + 0.5 * hs], dim=-1)

    def single_level_grid_anchors(self, base_anchors, featmap_size, stride=(16, 16), device='cuda'):
        fh, fw = featmap_size
        sx = torch.arange(0, fw, device=device) * stride[0]
        sy = torch.arange(0, fh, device=device) * stride[1]
        xx, yy = sx.repeat(fh), sy.view(-1, 1).repeat(1, fw).view(-1)
        shifts = torch.stack([xx, yy, xx, yy], dim=-1).type_as(base_anchors)
        return (base_anchors[None, :, :] + shifts[:, None, :]).view(-1, 4)

    def grid_anchors(self, featmap_sizes, device='cuda'):
        assert self.num_levels == len(featmap_sizes)
        return [self.single_level_grid_anchors(self.base_anchors[i].to(device), featmap_sizes[i], self.strides[i], device=device)
                for i in range(self.num_levels)]

    def valid_flags(self, featmap_sizes, pad_shape, device='cuda'):
        """anchor_generator.py:272-330: per level bool (H*W*A,), valid where the cell lies inside ceil(pad_shape / stride)."""
        assert self.num_levels == len(featmap_sizes)
        out = []
        for i, (fh, fw) in enumerate(featmap_sizes):
            sw, sh = self.strides[i]
            h, w = pad_shape[:2]
            vh, vw = min(int(np.ceil(h / sh)), fh), min(int(np.ceil(w / sw)), fw)
            vx = torch.zeros(fw, dtype=torch.bool, device=device)
            vy = torch.zeros(fh, dtype=torch.bool, device=device)
            vx[:vw] = 1
            vy[:vh] = 1
            v = vx.repeat(fh) & vy.view(-1, 1).repeat(1, fw).view(-1)
            A = self.num_base_anchors[i]
            out.append(v[:, None].expand(v.size(0), A).contiguous().view(-1))
        return out


class RPNProposals:
    """get_bboxes of the reference RPNHead (sigmoid classification, DeltaXYWHBBoxCoder) over ptb_rpn_proposals."""

    def __init__(self, anchor_generator, bbox_coder=None, test_cfg=None, use_sigmoid_cls=True):
        ag = dict(anchor_generator)
        if ag.pop('type', 'AnchorGenerator') != 'AnchorGenerator':
            raise NotImplementedError('only AnchorGenerator is implemented')
        self.anchor_generator = AnchorGenerator(**ag)
        bc = dict(bbox_coder or dict(type='DeltaXYWHBBoxCoder'))
        if bc.pop('type', 'DeltaXYWHBBoxCoder') != 'DeltaXYWHBBoxCoder':
            raise NotImplementedError('only DeltaXYWHBBoxCoder is implemented')
        if bc.get('add_ctr_clamp', False) or not bc.get('clip_border', True):
            raise NotImplementedError('DeltaXYWHBBoxCoder(add_ctr_clamp=True / clip_border=False)')
        self.means, self.stds = tuple(bc.get('target_means', (0., 0., 0., 0.))), tuple(bc.get('target_stds', (1., 1., 1., 1.)))
        self.wh_ratio_clip = 16 / 1000                       # DeltaXYWHBBoxCoder.decode default (delta_xywh_bbox_coder.py:88)
        if not use_sigmoid_cls:
            raise NotImplementedError('RPN softmax classification (loss_cls.use_sigmoid=False)')
        if len(set(self.anchor_generator.num_base_anchors)) != 1:
            raise NotImplementedError('levels with different numbers of base anchors')
        self.test_cfg = CfgNode(test_cfg) if test_cfg is not None else None
        self._base_dev = {}

    def _base(self, device):
        if device not in self._base_dev:
            self._base_dev[device] = torch.stack(self.anchor_generator.base_anchors).to(device).contiguous()
        return self._base_dev[device]

    @torch.no_grad()
    def get_bboxes(self, cls_scores, bbox_preds, img_metas, cfg=None, rescale=False, with_nms=True, return_levels=False):
        assert len(cls_scores) == len(bbox_preds) == self.anchor_generator.num_levels
        if not with_nms:
            raise NotImplementedError('with_nms=False')
        if not cls_scores[0].is_cuda:
            raise RuntimeError('RPNProposals (B200) runs on CUDA tensors only; there is no CPU fallback')
        cfg = CfgNode(cfg) if cfg is not None else self.test_cfg
        nms = dict(cfg.get('nms', dict(type='nms', iou_threshold=cfg.get('nms_thr', 0.7))))
        if nms.get('type', 'nms') != 'nms':
            raise NotImplementedError(f"rpn nms type {nms.get('type')}")
        max_per_img = cfg.get('max_per_img', cfg.get('max_num', cfg.get('nms_post', 1000)))   # older configs spell it max_num / nms_post
        dev = cls_scores[0].device
        img_hw = torch.tensor([[int(m['img_shape'][0]), int(m['img_shape'][1])] for m in img_metas], dtype=torch.int32).to(dev)
        cnt, det, lvl = ops.rpn_proposals([c.detach().float().contiguous() for c in cls_scores], [r.detach().float().contiguous() for r in bbox_preds],
                                          self._base(dev), self.anchor_generator.strides, img_hw, self.means, self.stds, self.wh_ratio_clip,
                                          cfg.get('nms_pre', -1), cfg.get('min_bbox_size', 0), nms.get('iou_threshold', 0.7), max_per_img)
        cnt = cnt.cpu().tolist()                      # ragged result lists, like the reference's per-image dets
        out = [det[b, :cnt[b]] for b in range(len(cnt))]
        if return_levels:
            return out, [lvl[b, :cnt[b]] for b in range(len(cnt))]
        return out
