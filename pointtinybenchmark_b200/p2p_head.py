"""P2PHead — host-side mirror of the reference's P2PNet-style head over the sm_100a kernels.

reference: TOV_mmdetection/mmdet/models/point/dense_heads/p2p_head.py:18-572 (P2PHead), HungarianAssignerV2
(core/bbox/assigners/hungarian_assigner.py:149-270), FocalLossCost/DisCostV2 (core/bbox/match_costs/match_cost.py),
multiclass_nms (core/post_processing/bbox_nms.py).  Same ctor kwargs / outputs / state_dict keys
(cls_convs.*, reg_convs.*, cls_out (conv3x3), reg_out (conv3x3)).

What runs where: towers and the two output convs = tcgen05 implicit GEMMs of libptb_b200.so at inference; under autograd the
towers use the tensor-core autograd function of layers.py (dgrad / wgrad / GroupNorm backward kernels) and the two narrow output
convs cuDNN fp32; decode, top-k, NMS / soft-NMS, cost matrix, the Hungarian matching (scipy's shortest-augmenting-path algorithm
restated as a one-CTA-per-image kernel, bit-identical assignments incl. ties: csrc/lsap_core.cuh) and the
focal / smooth-L1 losses = libptb_b200.so.  Nothing of the training step returns to the host except one (B,) status read.
"""
import numpy as np
import torch
import torch.nn as nn

from . import ops
from .layers import ConvModule, PackedWeightsMixin, bias_init_with_prob, normal_init_, tower, tc_enabled, _packed_tc
from .registry import CfgNode, register_head


def _iou_of(nms_cfg):
    """IoU threshold of an mmcv nms config: `iou_threshold` (mmcv >= 1.3) or its older spelling `iou_thr`."""
    v = nms_cfg.get('iou_threshold', None)
    v = nms_cfg.get('iou_thr', None) if v is None else v
    if v is None:
        raise KeyError("test_cfg.nms needs 'iou_threshold' (or 'iou_thr')")
    return float(v)


class _FocalSumFn(torch.autograd.Function):
    """sum_m,c sigmoid_focal(logits, labels) * weight[m]   (losses/focal_loss.py:11-56)"""

    @staticmethod
    def forward(ctx, logits, labels, weight, gamma, alpha):
        ctx.save_for_backward(logits, labels, weight)
        ctx.ga = (gamma, alpha)
        return ops.sigmoid_focal(logits, labels, weight, gamma, alpha)[0]

    @staticmethod
    def backward(ctx, g):
        logits, labels, weight = ctx.saved_tensors
        scale = g.reshape(1).float().contiguous()
        return ops.sigmoid_focal(logits, labels, weight, ctx.ga[0], ctx.ga[1], scale=scale, want_grad=True), None, None, None, None


class _SmoothL1SumFn(torch.autograd.Function):
    """sum smooth_l1((pred - target) * inv_norm, beta) * weight   (losses/smooth_l1_loss.py:25-31)"""

    @staticmethod
    def forward(ctx, pred, target, weight, inv_norm, beta):
        ctx.save_for_backward(pred, target, weight)
        ctx.nb = (inv_norm, beta)
        return ops.smooth_l1(pred, target, weight, inv_norm, beta)[0]

    @staticmethod
    def backward(ctx, g):
        pred, target, weight = ctx.saved_tensors
        scale = g.reshape(1).float().contiguous()
        return ops.smooth_l1(pred, target, weight, ctx.nb[0], ctx.nb[1], scale=scale, want_grad=True), None, None, None, None


@register_head
class P2PHead(PackedWeightsMixin, nn.Module):
    def __init__(self, num_classes, in_channels,
                 point_anchor=((-0.25, -0.25), (0.25, -0.25), (0.25, 0.25), (-0.25, 0.25)),
                 assign_before_pred=False, pts_gamma=100. / 8, reg_norm=1. / 8,
                 loss_cls=None, loss_reg=None, init_cfg=None,
                 feat_channels=256, stacked_convs=4, strides=(4, 8, 16, 32, 64),
                 conv_cfg=None, norm_cfg=None, conv_bias='auto', dcn_on_last_conv=False, loss_bbox=None,
                 train_cfg=None, test_cfg=None, **kwargs):
        super().__init__()
        if kwargs:
            raise TypeError(f'P2PHead: unexpected kwargs {sorted(kwargs)}')
        self.num_classes, self.in_channels, self.feat_channels = num_classes, in_channels, feat_channels
        self.stacked_convs, self.strides = stacked_convs, list(strides)
        self.point_anchor = torch.tensor(point_anchor, dtype=torch.float32).reshape(-1, 2)
        self.num_points = self.point_anchor.shape[0]
        self.assign_before_pred, self.pts_gamma, self.reg_norm = assign_before_pred, pts_gamma, reg_norm
        self.loss_cls_cfg = dict(type='CrossEntropyLoss', use_sigmoid=True, loss_weight=1.0)
        self.loss_cls_cfg.update(loss_cls or {})
        self.loss_reg_cfg = dict(type='MSELoss', loss_weight=2e-4)
        self.loss_reg_cfg.update(loss_reg or {})
        self.train_cfg = CfgNode(train_cfg) if train_cfg is not None else None
        self.test_cfg = CfgNode(test_cfg) if test_cfg is not None else None
        if len(self.strides) != 1:
            raise NotImplementedError('P2PHead (B200): one FPN level only (all configs2/*/p2p configs use strides=[s])')
        if not self.loss_cls_cfg.get('use_sigmoid', False):
            raise NotImplementedError('P2PHead (B200): softmax classification')
        self.num_cls_out = num_classes
        self.cls_convs, self.reg_convs = nn.ModuleList(), nn.ModuleList()
        for i in range(stacked_convs):
            chn = in_channels if i == 0 else feat_channels
            self.cls_convs.append(ConvModule(chn, feat_channels, 3, 1, 1, norm_cfg=norm_cfg, bias=conv_bias))
            self.reg_convs.append(ConvModule(chn, feat_channels, 3, 1, 1, norm_cfg=norm_cfg, bias=conv_bias))
        self.cls_out = nn.Conv2d(feat_channels, self.num_cls_out * self.num_points, 3, padding=1)
        self.reg_out = nn.Conv2d(feat_channels, self.num_points * 2, 3, padding=1)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                normal_init_(m, 0.01, 0.0)
        nn.init.constant_(self.cls_out.bias, bias_init_with_prob(0.01))
        if self.train_cfg is not None:
            a = dict(self.train_cfg.assigner)
            if a.get('type') != 'HungarianAssignerV2':
                raise NotImplementedError(f"assigner {a.get('type')}")
            cc, rc = a.get('cls_costs'), a.get('reg_costs')
            cc = cc[0] if isinstance(cc, (list, tuple)) else cc
            rc = rc[0] if isinstance(rc, (list, tuple)) else rc
            if cc.get('type') != 'FocalLossCost' or rc.get('type') != 'DisCostV2':
                raise NotImplementedError('only FocalLossCost + DisCostV2 match costs are implemented')
            self.assign = dict(w_cls=cc.get('weight', 1.0), alpha=cc.get('alpha', 0.25), gamma=cc.get('gamma', 2),
                               eps=cc.get('eps', 1e-12), w_dis=rc.get('weight', 1.0),
                               norm_wh=rc.get('norm_with_img_wh', True), p=rc.get('p', 1), topk_k=a.get('topk_k', 1))
            if self.assign['p'] != 1:
                raise NotImplementedError('DisCostV2 p != 1')
        self._init_packed_hooks()
        self.check_assign_status = True      # read the (B,) status of the matching kernel each step (scipy's ValueErrors)

    # ------------------------------------------------------------------------------------------------
    def forward(self, feats):
        """p2p_head.py:104-123.  Inference: towers AND the two conv3x3 output layers run on the tcgen05 kernel (fp16 two-term
        split, fp32-level accuracy); with autograd recording the towers use the tensor-core autograd function of layers.py and the
        two output convs cuDNN fp32."""
        cls_outs, pts_outs = [], []
        for x in feats:
            if tc_enabled(x, self.cls_convs, self.reg_convs, self.cls_out, self.reg_out) and self.feat_channels == 256 \
                    and self.in_channels % 32 == 0:
                info = {}
                pc = tower(self.cls_convs, x, info, want='f16pair')
                pr = tower(self.reg_convs, x, info, want='f16pair') if pc is not None else None
                if pc is not None and pr is not None:
                    self.last_tower_backend = info.get('backend')
                    nc, nr = self.cls_out.out_channels, self.reg_out.out_channels
                    yc = ops.conv_tc_f16(pc[0], pc[1], _packed_tc(self.cls_out, 9, 'conv'), 9, nc, bias=self.cls_out.bias.detach())
                    yr = ops.conv_tc_f16(pr[0], pr[1], _packed_tc(self.reg_out, 9, 'conv'), 9, nr, bias=self.reg_out.bias.detach())
                    cls_outs.append(yc[..., :nc].permute(0, 3, 1, 2))
                    pts_outs.append(yr[..., :nr].permute(0, 3, 1, 2))
                    continue
            # autograd path: the towers run layers._TowerTCFn (tensor-core forward + dgrad + wgrad + GroupNorm backward) for the
            # shipped geometry; the two narrow output convs (256 -> C / 2k) stay cuDNN fp32, never TF32 (1e-4 logits)
            info = {}
            fc, fr = tower(self.cls_convs, x, info), tower(self.reg_convs, x, info)
            self.last_tower_backend = info.get('backend')
            with torch.backends.cudnn.flags(enabled=torch.backends.cudnn.enabled, benchmark=torch.backends.cudnn.benchmark,
                                            deterministic=torch.backends.cudnn.deterministic, allow_tf32=False):
                cls_outs.append(self.cls_out(fc))
                pts_outs.append(self.reg_out(fr))
        return cls_outs, pts_outs

    def forward_train(self, x, img_metas, gt_bboxes, gt_labels=None, gt_bboxes_ignore=None, proposal_cfg=None, **kwargs):
        outs = self(x)
        return self.loss(*outs, gt_bboxes, gt_labels, img_metas, gt_bboxes_ignore=gt_bboxes_ignore)

    def simple_test(self, feats, img_metas, rescale=False, **kwargs):
        outs = self.forward(feats)
        return self.get_bboxes(*outs, img_metas, rescale=rescale)

    # ------------------------------------------------------------------------------------------------
    def _grid(self, H, W, device):
        s = float(self.strides[0])
        xx = (torch.arange(0., W, device=device) * s).repeat(H)
        yy = (torch.arange(0., H, device=device) * s).view(-1, 1).repeat(1, W).view(-1)
        return xx, yy

    def get_pred_points(self, cls_out, pts_out, img_metas):
        """p2p_head.py:125-170 (single level): differentiable torch elementwise ops on channels-last views."""
        B, _, H, W = cls_out.shape
        k, C, s = self.num_points, self.num_cls_out, float(self.strides[0])
        dev = cls_out.device
        cls = ops.to_nhwc(cls_out).reshape(B, H * W * k, C)
        reg = ops.to_nhwc(pts_out).reshape(B, H * W, k, 2)
        xx, yy = self._grid(H, W, dev)
        anchor = torch.stack([xx, yy], -1)[None, :, None, :] + (self.point_anchor.to(dev) * s)[None, None]
        anchor = anchor.expand(B, H * W, k, 2)
        pred = anchor + reg * self.pts_gamma * s
        vflag = []
        for m in img_metas:
            ph, pw = m['pad_shape'][:2]
            vh, vw = min(int(np.ceil(ph / s)), H), min(int(np.ceil(pw / s)), W)
            v = torch.zeros(H, W, dtype=torch.bool)
            v[:vh, :vw] = True
            vflag.append(v.reshape(-1))
        self._valid_host = vflag                       # host copy: the assignment sizes its launches without a device sync
        valid = torch.stack(vflag).to(dev)[:, :, None].expand(B, H * W, k).reshape(B, -1)
        return anchor.reshape(B, -1, 2), pred.reshape(B, -1, 2), valid, cls

    def loss(self, cls_outs, pts_outs, gt_bboxes, gt_labels, img_metas, gt_bboxes_ignore=None):
        """p2p_head.py:172-248 -> dict(loss_cls=[B], loss_pts=[B])."""
        cls_out, pts_out = cls_outs[0], pts_outs[0]
        if not cls_out.is_cuda:
            raise RuntimeError('P2PHead (B200) runs on CUDA tensors only; there is no CPU fallback')
        for gb in gt_bboxes:                       # p2p_head.py:183-184: the reference refuses images without a GT point
            assert len(gb) > 0, gt_bboxes
        dev = cls_out.device
        anchor, pred, valid, cls = self.get_pred_points(cls_out, pts_out, img_metas)
        B, Q, C = cls.shape
        s = float(self.strides[0])
        prop = (anchor if self.assign_before_pred else pred).detach().contiguous()
        a = self.assign
        # ---- cost matrices and the Hungarian matching on the GPU: no cost.cpu(), no scipy (hungarian_assigner.py:229-270)
        key = (tuple(cls_out.shape[-2:]), tuple(tuple(m['pad_shape'][:2]) for m in img_metas), str(dev))
        if getattr(self, '_ridx_cache', (None,))[0] != key:           # valid-row lists depend on the map size and pad shapes only
            valid_host = torch.stack(self._valid_host)[:, :, None].expand(B, valid.shape[1] // self.num_points, self.num_points).reshape(B, -1)
            ridx_l = [torch.nonzero(valid_host[b]).squeeze(1).int() for b in range(B)]
            n_valid = [int(r.shape[0]) for r in ridx_l]
            r_off = np.concatenate([[0], np.cumsum(n_valid)]).astype(np.int64)
            r_flat = torch.cat(ridx_l).to(dev) if r_off[-1] > 0 else torch.zeros(1, dtype=torch.int32, device=dev)
            self._ridx_cache = (key, n_valid, r_off, r_flat)
        _, n_valid, ridx_off, ridx_flat = self._ridx_cache
        shapes = [(n_valid[b], int(gt_labels[b].shape[0])) for b in range(B)]
        cost_off = np.concatenate([[0], np.cumsum([sh[0] * sh[1] for sh in shapes])]).astype(np.int64)
        cost_flat = torch.empty(max(int(cost_off[-1]), 1), dtype=torch.float32, device=dev)
        gpts_l = []
        for b in range(B):
            gpts = ((gt_bboxes[b][:, :2] + gt_bboxes[b][:, 2:]) / 2).to(dev).float().contiguous()
            gpts_l.append(gpts)
            if shapes[b][0] == 0 or shapes[b][1] == 0:
                continue
            fx, fy = (img_metas[b]['img_shape'][1], img_metas[b]['img_shape'][0]) if a['norm_wh'] else (1.0, 1.0)
            ops.p2p_cost_matrix(cls[b].detach().contiguous(), prop[b], ridx_flat[ridx_off[b]:ridx_off[b + 1]], gpts,
                                gt_labels[b].to(dev).int().contiguous(), a['w_cls'], a['alpha'], a['gamma'], a['eps'], a['w_dis'], fx, fy,
                                out=cost_flat[cost_off[b]:cost_off[b + 1]])
        gi_all = torch.zeros((B, Q), dtype=torch.int64, device=dev)
        status = ops.hungarian_v2_batch(cost_flat, shapes, a['topk_k'], gi_all, [b * Q for b in range(B)], ridx_flat, ridx_off[:-1])
        if self.check_assign_status:                          # one (B,) int32 read per batch: scipy's two ValueErrors
            st = status.cpu()
            if int(st.max()) != 0:
                bad = int(torch.nonzero(st)[0])
                raise ValueError({1: 'cost matrix is infeasible', 2: 'matrix contains invalid numeric entries'}.get(
                    int(st[bad]), f'hungarian kernel status {int(st[bad])}') + f' (image {bad})')
        self._last_assign = dict(gt_inds=gi_all, status=status)
        labels_l, lw_l, gp_l, pw_l = [], [], [], []
        neg_w = self.train_cfg.get('neg_weight', 1.0)
        pos_w = self.train_cfg.get('pos_weight', 1.0)
        for b in range(B):
            gi = gi_all[b]
            vb = valid[b]
            pos = gi > 0
            gl = gt_labels[b].to(dev)
            gpts = gpts_l[b]
            if gl.shape[0] == 0:                                 # no GT: everything background (hungarian_assigner.py:214-219)
                gl = torch.zeros(1, dtype=torch.long, device=dev)
                gpts = torch.zeros(1, 2, device=dev)
            labels = torch.where(pos, gl[(gi - 1).clamp(min=0)], torch.full_like(gi, self.num_classes))
            labels = torch.where(vb, labels, torch.zeros_like(labels))               # unmap(fill=0)
            lw = torch.where(pos, torch.full((Q,), float(pos_w), device=dev),
                             torch.full((Q,), 1.0 if neg_w <= 0 else float(neg_w), device=dev)) * vb.float()
            gp = torch.where(pos[:, None], gpts[(gi - 1).clamp(min=0)], torch.zeros(Q, 2, device=dev))
            pw = pos.float()[:, None].expand(Q, 2).contiguous()
            labels_l.append(labels); lw_l.append(lw.contiguous()); gp_l.append(gp.contiguous()); pw_l.append(pw)
        num_total_pos = sum([(p[:, 0] > 0).sum() for p in pw_l]).float()
        gamma, alpha = self.loss_cls_cfg.get('gamma', 2.0), self.loss_cls_cfg.get('alpha', 0.25)
        if self.loss_cls_cfg['type'] != 'FocalLoss' or self.loss_reg_cfg['type'] != 'SmoothL1Loss':
            raise NotImplementedError('P2PHead (B200): loss_cls must be FocalLoss and loss_reg SmoothL1Loss')
        loss_cls, loss_pts = [], []
        for b in range(B):
            lc = _FocalSumFn.apply(cls[b].contiguous(), labels_l[b], lw_l[b], gamma, alpha)
            loss_cls.append(self.loss_cls_cfg.get('loss_weight', 1.0) * lc / num_total_pos)
            lp = _SmoothL1SumFn.apply(pred[b].contiguous(), gp_l[b], pw_l[b], 1.0 / (s * self.reg_norm),
                                      self.loss_reg_cfg.get('beta', 1.0))
            loss_pts.append(self.loss_reg_cfg.get('loss_weight', 1.0) * lp / num_total_pos)
        self._last_targets = dict(labels=labels_l, label_weights=lw_l, gt_pts=gp_l, pts_weights=pw_l)
        return dict(loss_cls=loss_cls, loss_pts=loss_pts)

    # ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def get_bboxes(self, cls_outs, pts_outs, img_metas, cfg=None, rescale=False, with_nms=True, return_all=False):
        """p2p_head.py:330-423: per image (pseudo boxes (m,5), labels (m,))."""
        cfg = CfgNode(cfg) if cfg is not None else self.test_cfg
        cls_out, pts_out = cls_outs[0], pts_outs[0]
        if not cls_out.is_cuda:
            raise RuntimeError('P2PHead (B200) runs on CUDA tensors only; there is no CPU fallback')
        if not with_nms:
            raise NotImplementedError('with_nms=False')
        dev = cls_out.device
        B = cls_out.shape[0]
        img_hw = torch.tensor(np.array([m['img_shape'][:2] for m in img_metas], dtype=np.int32), device=dev)
        scale_xy = None
        if rescale:
            scale_xy = torch.tensor(np.array([m['scale_factor'][:2] for m in img_metas], dtype=np.float32), device=dev)
        cmap, rmap = ops.to_nhwc(cls_out).contiguous(), ops.to_nhwc(pts_out).contiguous()
        idx, pts, scores = ops.p2p_decode_topk(cmap, rmap, self.num_cls_out, self.num_points, self.point_anchor.to(dev),
                                               self.strides[0], self.pts_gamma, img_hw, cfg.get('nms_pre', -1), scale_xy)
        wh = cfg.get('pseudo_wh', (16, 16))
        nms = cfg.get('nms')
        if nms.get('type', 'nms') == 'soft_nms':       # batched_nms dispatches on nms_cfg['type'] (mmcv/ops/nms.py)
            cnt, det, lab, keep, cc = ops.multiclass_soft_nms(pts, scores, wh, cfg.get('score_thr'), nms.get('iou_threshold', 0.3),
                                                              cfg.get('max_per_img'), nms.get('sigma', 0.5),
                                                              nms.get('min_score', 1e-3), nms.get('method', 'linear'))
        elif nms.get('type', 'nms') == 'nms':
            cnt, det, lab, keep, cc = ops.multiclass_nms(pts, scores, wh, cfg.get('score_thr'), _iou_of(nms),
                                                         cfg.get('max_per_img'))
        else:
            raise NotImplementedError(f"nms type {nms.get('type')}")
        cnt_h = cnt.cpu().tolist()
        res = []
        for b in range(B):
            m = cnt_h[b]
            d = det[b, :m]
            cxcy = torch.stack([(d[:, 0] + d[:, 2]) / 2, (d[:, 1] + d[:, 3]) / 2], -1)      # bbox_xyxy_to_cxcywh
            half = cxcy.new_tensor(wh) / 2
            res.append((torch.cat([cxcy - half, cxcy + half, d[:, 4:5]], -1), lab[b, :m].long()))
        if return_all:
            return res, dict(topk_idx=idx, pts=pts, scores=scores, keep=keep, count=cnt, cand_count=cc)
        return res

    # ------------------------------------------------------------------------------------------------
    @staticmethod
    def bbox_mapping_back(bboxes, img_shape, scale_factor, flip, flip_direction, tile_offset=None):
        """core/bbox/transforms.py:62-85 (incl. the reference's tile_offset extension)."""
        b = bboxes
        if flip:
            f = b.clone()
            if flip_direction in ('horizontal', 'diagonal'):
                f[..., 0::4] = img_shape[1] - b[..., 2::4]
                f[..., 2::4] = img_shape[1] - b[..., 0::4]
            if flip_direction in ('vertical', 'diagonal'):
                f[..., 1::4] = img_shape[0] - b[..., 3::4]
                f[..., 3::4] = img_shape[0] - b[..., 1::4]
            b = f
        b = b.view(-1, 4) / b.new_tensor(scale_factor)
        if tile_offset is not None:
            dx, dy = tile_offset
            b[:, [0, 2]] += dx
            b[:, [1, 3]] += dy
        return b.view(bboxes.shape)

    @torch.no_grad()
    def aug_test_bboxes(self, feats, img_metas, rescale=False):
        """test-time-aug / cropped-tile merge (p2p_head.py:487-572, dense_test_mixins.py:173-204): per aug run get_bboxes
        (with NMS), map the kept boxes back, then a SECOND multiclass NMS over the union (ptb_multiclass_nms_boxes)."""
        aug_bboxes, aug_scores = [], []
        for x, img_meta in zip(feats, img_metas):
            outs = self.forward(x)
            boxes5, labels = self.get_bboxes(*outs, img_meta, cfg=self.test_cfg, rescale=False, with_nms=True)[0]
            sc = boxes5.new_zeros((boxes5.shape[0], self.num_classes))
            sc[torch.arange(boxes5.shape[0], device=boxes5.device), labels] = boxes5[:, 4]
            m = img_meta[0]
            aug_bboxes.append(self.bbox_mapping_back(boxes5[:, :4], m['img_shape'], m['scale_factor'], m.get('flip', False),
                                                     m.get('flip_direction', 'horizontal'), m.get('tile_offset', None)))
            aug_scores.append(sc)
        boxes = torch.cat(aug_bboxes).contiguous()
        scores = torch.cat(aug_scores).contiguous()
        cfg = self.test_cfg
        if boxes.shape[0] == 0:
            return [(boxes.new_zeros((0, 5)), boxes.new_zeros((0,), dtype=torch.long))]
        nms = cfg.get('nms')
        if nms.get('type', 'nms') == 'soft_nms':
            cnt, det, lab, _, _ = ops.multiclass_soft_nms(boxes[None], scores[None], None, cfg.get('score_thr'),
                                                          nms.get('iou_threshold', 0.3), cfg.get('max_per_img'), nms.get('sigma', 0.5),
                                                          nms.get('min_score', 1e-3), nms.get('method', 'linear'))
        else:
            cnt, det, lab, _, _ = ops.multiclass_nms_boxes(boxes[None], scores[None], cfg.get('score_thr'),
                                                           _iou_of(nms), cfg.get('max_per_img'))
        n = int(cnt[0])
        d = det[0, :n].clone()
        if not rescale:
            d[:, :4] *= d.new_tensor(img_metas[0][0]['scale_factor'])
        return [(d, lab[0, :n].long())]
