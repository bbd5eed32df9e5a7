"""ORACLE (test infrastructure, NOT product code) — CPU restatement of the reference CPR head.

Restates, in plain functional torch-CPU code, the algorithm of
  /root/reference/TOV_mmdetection/mmdet/models/point/dense_heads/cpr_head.py  (CPRHead, generators,
  PointExtractor, PointRefiner) and mmdet/models/losses/multi_instance_learning_loss.py (MILLoss).
Every function cites the reference file:line it follows ("ref:" = cpr_head.py unless a file is named).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package; the product package `pointtinybenchmark_b200` never does.

PINNED: oracle/make_golden.py executes the real reference (via oracle/_mmcv_stub.py) in the build
container on seeded inputs and asserts this restatement reproduces it (bit-exact for every integer/bool
output and for the floats, since both sides call the same ATen CPU kernels in the same order); the
resulting vectors are committed under tests/golden/.

The float ops deliberately go through the same ATen CPU kernels the reference uses (F.grid_sample,
torch.cdist, F.linear, softmax, ...) so the oracle defines the same rounding as the reference's CPU path.
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# config (mirrors the ctor kwargs of CPRHead, ref:904-981; defaults = configs2/_base_/models/cpr/
# coarse_point_refine_r50_fpns4_1x.py:26-69 merged with COCO/coarsepointv2/...coco400.py:80-91)
# ----------------------------------------------------------------------------------------------


def default_cfg(**over):
    cfg = dict(
        num_classes=80, in_channels=256, feat_channels=256, stacked_convs=4, stride=8, gn_groups=32,
        pos_radius=8, neg_radius=8, neg_class_wise=True,
        pos_generator='circle',            # 'circle' (CirclePtFeatGenerator) | 'grid_circles' (GridCirclesPtFeatGenerator)
        grid_max_pos_num=-1,               # GridPtFeatGenerator.max_pos_num (<= 0: 2*(2*radius)**2, ref:440-444)
        start_angle=0, base_num_point=8, same_num_all_radius=False, append_center=True,
        mil_loss_weight=0.25, mil_eps=1e-6,
        with_neg=True, neg_loss_weight=0.75, refine_bag_policy='only_refine_bag',
        with_gt_loss=True, gt_loss_type='gt_refine', gt_loss_weight=0.125, with_mil_loss=True,
        prob_cls_type='sigmoid', normed_sigmoid_p=1, binary_ins=False, num_cls_fcs=0, mil_loss_type='gfocal_loss',
        gt_alpha=0.5, merge_th=0.1, refine_th=0.1, classify_filter=True, nearest_filter=True,
        return_score_type='mean',
    )
    cfg.update(over)
    return cfg


# ----------------------------------------------------------------------------------------------
# conv towers  (ref:983-995 _init_layers, ref:1033-1043 forward_single)
# ----------------------------------------------------------------------------------------------
def tower_forward(x, weights, cfg, prefix='cls_convs'):
    """4 x [conv3x3 (no bias) -> GroupNorm(32) -> ReLU]; parameter names follow mmcv ConvModule."""
    for i in range(cfg['stacked_convs']):
        x = F.conv2d(x, weights[f'{prefix}.{i}.conv.weight'], None, 1, 1)
        x = F.group_norm(x, cfg['gn_groups'], weights[f'{prefix}.{i}.gn.weight'], weights[f'{prefix}.{i}.gn.bias'])
        x = F.relu(x)
    return x


# ----------------------------------------------------------------------------------------------
# bag geometry
# ----------------------------------------------------------------------------------------------
def circle_offsets(radius, stride, start_angle=0, base_num_point=8, same_num_all_radius=False):
    """ring offsets, ref:484-491 (CirclePtFeatGenerator.get_point_neighbours).
    ring i: r=(i+1)*stride, m = 8*(i+1) points at angle (j/m*360+start)/360*pi*2."""
    out = []
    for i in range(radius):
        r = (i + 1) * stride
        m = base_num_point if same_num_all_radius else base_num_point * (i + 1)
        ang = torch.arange(m).float() / m * 360 + start_angle
        ang = ang / 360 * np.pi * 2
        out.append(torch.stack([r * torch.cos(ang), r * torch.sin(ang)], dim=-1))
    return torch.cat(out)


def circle_bag_points(centers, stride, cfg, radius):
    """(G,2) -> (G,K,2): rings + centre LAST (ref:492-497)."""
    off = circle_offsets(radius, stride, cfg['start_angle'], cfg['base_num_point'], cfg['same_num_all_radius'])
    pts = off.unsqueeze(0) + centers.reshape(-1, 1, 2)
    if cfg['append_center']:
        pts = torch.cat([pts, centers.unsqueeze(1)], dim=1)
    return pts


def point_valid(pts, valid_h, valid_w):
    """ref:172-180: 0<=x<valid_w & 0<=y<valid_h (image coords vs pad_shape)."""
    return (0 <= pts[..., 0]) & (pts[..., 0] < valid_w) & (0 <= pts[..., 1]) & (pts[..., 1] < valid_h)


def sample_point_feat(feat, pts, stride):
    """ref:182-199 extract_point_feat + ref:73-93 grid_sample (align_corners=False, border padding).
    feat (1,C,H,W), pts (..., K, 2) image coords -> (..., K, C)."""
    s = pts.shape[:-2]
    p = pts.flatten(0, -3).unsqueeze(0) / stride
    h, w = feat.shape[2:]
    wh = feat.new_tensor([w, h])
    grid = (2 * p + 1) / wh - 1
    out = F.grid_sample(feat, grid, align_corners=False, padding_mode='border').permute(0, 2, 3, 1)[0]
    return out.reshape(*s, out.shape[-2], out.shape[-1])


def anchor_points(h, w, valid_h, valid_w, stride):
    """ref:240-244 AnchorPtFeatGenerator.anchor_points: (j*s+s/2, i*s+s/2), valid vs pad_shape."""
    y, x = torch.meshgrid(torch.arange(h), torch.arange(w), indexing='ij')
    pts = torch.stack([x, y], dim=-1) * stride + stride / 2
    return pts, point_valid(pts, valid_h, valid_w)


def out_circle_neg_mask(grid_pts, grid_valid, centers, labels, stride, radius, num_classes, class_wise=True):
    """ref:254-290 OutCirclePtFeatGenerator.generate.
    grid_pts (HW,2), grid_valid (HW,), centers (n,R,2), labels (n,) -> bool (HW,num_classes).
    NOTE torch.cdist takes the matmul formulation for >25 rows (ATen cdist_impl), so the mask carries
    that rounding; the CUDA kernel reproduces the same operation order."""
    valid = grid_valid.reshape(-1, 1).repeat(1, num_classes)
    if class_wise:
        lab = labels.tolist()
        groups = OrderedDict()
        for l, c in zip(lab, centers):
            groups.setdefault(l, []).append(c)
        for l, cs in groups.items():
            c = torch.stack(cs).flatten(0, 1)
            dist = torch.cdist(grid_pts[..., :2], c, p=2)
            chosen = dist.min(dim=1)[0] >= stride * radius
            valid[..., l] = (valid[..., l].float() * chosen.float()).bool()
        return valid
    c = centers.flatten(0, 1)
    dist = torch.cdist(grid_pts[..., :2], c, 2)
    chosen = dist.min(dim=1)[0] >= stride * radius
    return (valid.float() * chosen[..., None].float()).bool()


def grid_circles_chosen(grid_pts_hw, centers, stride, radius):
    """ref:418-433 GridCirclesPtFeatGenerator.get_chosen_neighbours: (H,W,2),(n,R,2) -> bool (n,H,W)."""
    H, W, _ = grid_pts_hw.shape
    n, R, _ = centers.shape
    dis = torch.norm(grid_pts_hw.reshape(1, H, W, 1, 2) - centers.reshape(n, 1, 1, R, 2), p=2, dim=-1)
    return torch.any(dis <= radius * stride, dim=-1)


def grid_circles_max_pos_num(radius, max_pos_num=-1):
    """ref:440-444 GridCirclesPtFeatGenerator.get_max_pos_num."""
    return 2 * (2 * radius) ** 2 if max_pos_num <= 0 else max_pos_num


def grid_circles_bag(feat, centers, valid_h, valid_w, stride, radius, max_pos_num=-1, keep_feats=True):
    """ref:296-350 GridPtFeatGenerator.generate with GridCircles neighbours, one image.
    feat (1,C,H,W), centers (n,R,2) ->
      pts   (n,1,Kt,3)  chosen grid-cell centres in row-major order, zero padded to max_pos_num+R slots, then the R centres
                        (refine order flipped) with the stride column;  Kt = max_pos_num + 2R (the reference adds R twice)
      valid (n,1,Kt,1)  True for filled slots and for the centres (NOT tested against pad_shape: ref:322 starts from ones)
      feats (n,1,Kt,C)  exact copies of the map cells; bilinear samples for the centres
      chosens (n,H,W)   the neighbour mask itself."""
    _, C, H, W = feat.shape
    g, _ = anchor_points(H, W, valid_h, valid_w, stride)                  # (H,W,2)
    g = g.float()
    g3 = torch.cat([g, torch.full(g.shape[:-1] + (1,), float(stride))], dim=-1)
    chosens = grid_circles_chosen(g, centers, stride, radius)
    n, R, _ = centers.shape
    slots = int(grid_circles_max_pos_num(radius, max_pos_num)) + R
    pos_pts = torch.zeros(n, slots, 3)
    valid = torch.ones(n, slots, 1, dtype=torch.bool)
    fmap = feat.permute(0, 2, 3, 1).squeeze(0)                           # (H,W,C)
    bag = torch.zeros(n, slots, C) if keep_feats else None
    for i, ch in enumerate(chosens):
        p = g3[ch].reshape(-1, 3)
        pos_pts[i, :len(p)] = p                                          # raises like the reference when len(p) > slots
        valid[i, len(p):] = False
        if keep_feats:
            bag[i, :len(p)] = fmap[ch].reshape(-1, C)
    c3 = torch.cat([centers, torch.full(centers.shape[:-1] + (1,), float(stride))], dim=-1)
    pos_pts = torch.cat([pos_pts, c3.flip(dims=(1,))], dim=1).unsqueeze(1)
    valid = torch.cat([valid, torch.ones(n, R, 1, dtype=torch.bool)], dim=1).unsqueeze(1)
    if keep_feats:
        cf = sample_point_feat(feat, centers, stride)                    # (n,R,C)
        bag = torch.cat([bag, cf.flip(dims=(1,))], dim=1).unsqueeze(1)
    return pos_pts, valid, bag, chosens


# ----------------------------------------------------------------------------------------------
# extraction over a batch (PointExtractor.extract ref:638-662, per-image loop ref:152-160)
# ----------------------------------------------------------------------------------------------
def pseudo_bbox_to_center(gt_bboxes):
    """ref:1293-1301"""
    return [(b[:, :2] + b[:, 2:]) / 2 for b in gt_bboxes]


def center_to_pseudo_bbox(centers, pseudo_wh=(16, 16)):
    """ref:1303-1309"""
    wh = centers[0].new_tensor(pseudo_wh)
    return [torch.cat([c - wh / 2, c + wh / 2], dim=-1) for c in centers]


def extract(cls_feat, gt_r_points, gt_labels, img_metas, cfg, keep_feats=True):
    """positive bags + negative grid for every image; returns concatenated tensors + per-image lengths.
      pos_pts (G,R,K,3) [x,y,stride], pos_valid (G,R,K,1) bool, pos_feats (G,R,K,C)
      neg_pts (B*HW,3),  neg_valid (B*HW,Ncls) bool,     neg_feats (B*HW,C)"""
    stride = cfg['stride']
    B, C, H, W = cls_feat.shape
    pos_pts, pos_valid, pos_feats, neg_pts, neg_valid, neg_feats = [], [], [], [], [], []
    for b in range(B):
        centers = gt_r_points[b]                      # (n,R,2)
        n, R, _ = centers.shape
        ph, pw = img_metas[b]['pad_shape'][:2]
        feat = cls_feat[b:b + 1]
        if cfg.get('pos_generator', 'circle') == 'grid_circles':
            pts3, valid4, bag, _ = grid_circles_bag(feat, centers, ph, pw, stride, cfg['pos_radius'],
                                                    cfg.get('grid_max_pos_num', -1), keep_feats)
            if keep_feats:
                pos_feats.append(bag)
            pos_pts.append(pts3)
            pos_valid.append(valid4)
        else:
            pts = circle_bag_points(centers.flatten(0, 1), stride, cfg, cfg['pos_radius']).reshape(n, R, -1, 2)
            valid = point_valid(pts, ph, pw)
            if keep_feats:
                pos_feats.append(sample_point_feat(feat, pts, stride))
            pts3 = torch.cat([pts, torch.full(pts.shape[:-1] + (1,), float(stride))], dim=-1)   # ref:201-204
            pos_pts.append(pts3)
            pos_valid.append(valid[..., None])
        g, gv = anchor_points(H, W, ph, pw, stride)
        g3 = torch.cat([g, torch.full(g.shape[:-1] + (1,), float(stride))], dim=-1).flatten(0, -2)
        nv = out_circle_neg_mask(g.flatten(0, -2), gv.flatten(), centers, gt_labels[b], stride,
                                 cfg['neg_radius'], cfg['num_classes'], cfg['neg_class_wise'])
        neg_pts.append(g3)
        neg_valid.append(nv)
        if keep_feats:
            neg_feats.append(feat.permute(0, 2, 3, 1).squeeze(0).flatten(0, -2))
    out = dict(pos_pts=torch.cat(pos_pts), pos_valid=torch.cat(pos_valid),
               neg_pts=torch.cat(neg_pts), neg_valid=torch.cat(neg_valid),
               pos_len=[len(p) for p in pos_pts], neg_len=[len(p) for p in neg_pts])
    if keep_feats:
        out['pos_feats'] = torch.cat(pos_feats)
        out['neg_feats'] = torch.cat(neg_feats)
    return out


def pts_outs(feats, weights, name, num_fcs=0):
    """ref:1045-1078 get_pts_outs: num_cls_fcs x (Linear + ReLU) (`cls_fcs`; with ins_share_head_feat the instance head sees the same
    post-FC features, ref:1065), then Linear(-> Ncls [x2 for binary_ins]) on every sampled point."""
    s = feats.shape
    x = feats.flatten(0, -2)
    for i in range(num_fcs):
        x = F.relu(F.linear(x, weights[f'cls_fcs.{i}.weight'], weights[f'cls_fcs.{i}.bias']))
    return F.linear(x, weights[f'{name}.weight'], weights[f'{name}.bias']).reshape(*s[:-1], -1)


def cls_prob(cls_out, cfg):
    """ref:1080-1099 (only 'sigmoid' is used by the shipped configs)."""
    t = cfg['prob_cls_type']
    if t == 'sigmoid':
        return cls_out.sigmoid()
    shape = cls_out.shape[:-1]
    x = cls_out.reshape(*shape, cfg['num_classes'], -1)          # (..., C, 1): the reference reduces over dim -2 (other ATen kernel than dim -1)
    if t == 'softmax':
        return x.softmax(dim=-2).reshape(*shape, -1)
    if t == 'normed_sigmoid':
        return F.normalize(x.sigmoid(), p=cfg.get('normed_sigmoid_p', 1), dim=-2).reshape(*shape, -1)
    raise ValueError(t)


# ----------------------------------------------------------------------------------------------
# losses
# ----------------------------------------------------------------------------------------------
def gfocal_loss(p, q, w, eps=1e-6):
    """multi_instance_learning_loss.py:148-151"""
    l1 = (p - q) ** 2
    l2 = q * (p + eps).log() + (1 - q) * (1 - p + eps).log()
    return -(l1 * l2 * w).sum(dim=-1)


def mil_bag_prob(bag_cls_prob, bag_ins_outs, valid):
    """multi_instance_learning_loss.py:166-171: softmax over the bag dim x valid, L1-normalise, sum."""
    B, N, C = bag_cls_prob.shape
    prob_ins = bag_ins_outs.reshape(B, N, C, -1).softmax(dim=1) * valid.unsqueeze(-1)
    prob_ins = F.normalize(prob_ins, dim=1, p=1)
    return (bag_cls_prob.unsqueeze(-1) * prob_ins).sum(dim=1)[..., 0]


def mil_loss(bag_cls_prob, bag_ins_outs, labels, valid, loss_weight=1.0, eps=1e-6, binary_ins=False):
    """multi_instance_learning_loss.py:153-203 (gfocal_loss; binary_ins doubles the instance head: a positive and a negative bag
    probability per class, the negative one trained towards 0, :179-186).  returns loss, acc(top1 %), num_sample, prob(B,C)."""
    B, N, C = bag_cls_prob.shape
    if binary_ins:
        prob_ins = bag_ins_outs.reshape(B, N, C, 2).softmax(dim=1) * valid.unsqueeze(-1)
        prob_ins = F.normalize(prob_ins, dim=1, p=1)
        prob2 = (bag_cls_prob.unsqueeze(-1) * prob_ins).sum(dim=1)                      # (B, C, 2)
        pred_label = prob2[..., 0].topk(1, dim=1)[1][:, 0]
        acc = (pred_label == labels).float().sum(0, keepdim=True) * (100.0 / max(B, 1))
        label_weights = (valid.sum(dim=1) > 0).float()
        onehot = torch.zeros(B, C)
        onehot[torch.arange(B), labels] = 1
        num_sample = max(torch.sum(label_weights.sum(dim=-1) > 0).float().item(), 1.)
        prob = torch.cat([prob2[..., 0], prob2[..., 1]])
        loss = gfocal_loss(prob, torch.cat([onehot, torch.zeros_like(onehot)]), torch.cat([label_weights, label_weights]), eps)
        return loss.sum() / num_sample * loss_weight, acc, num_sample, prob2[..., 0]
    prob = mil_bag_prob(bag_cls_prob, bag_ins_outs, valid)
    pred_label = prob.topk(1, dim=1)[1][:, 0]
    acc = (pred_label == labels).float().sum(0, keepdim=True) * (100.0 / max(B, 1))
    label_weights = (valid.sum(dim=1) > 0).float()
    onehot = torch.zeros(B, C)
    onehot[torch.arange(B), labels] = 1
    num_sample = max(torch.sum(label_weights.sum(dim=-1) > 0).float().item(), 1.)
    loss = gfocal_loss(prob, onehot, label_weights, eps)
    loss = loss.sum() / num_sample * loss_weight
    return loss, acc, num_sample, prob


def cpr_loss(cls_feat, weights, gt_bboxes, gt_labels, img_metas, cfg, return_all=False, gt_weights=None):
    """CPRHead.loss + loss0 (ref:1101-1117, 1131-1229) for ins_share_head_feat=True, R=num_refine>=1."""
    gt_points = pseudo_bbox_to_center(gt_bboxes)
    gt_r_points = [p.reshape(len(l), -1, *p.shape[1:]) for p, l in zip(gt_points, gt_labels)]
    ex = extract(cls_feat, gt_r_points, gt_labels, img_metas, cfg)
    nf = cfg.get('num_cls_fcs', 0)
    pos_cls = pts_outs(ex['pos_feats'], weights, 'cls_out', nf)
    pos_ins = pts_outs(ex['pos_feats'], weights, 'ins_out', nf)
    neg_cls = pts_outs(ex['neg_feats'], weights, 'cls_out', nf)
    labels_all = torch.cat(gt_labels)
    gt_weights = torch.ones(len(labels_all)) if gt_weights is None else torch.cat(list(gt_weights)).float()      # ref:1108-1114
    pos_pts, pos_valid, neg_valid = ex['pos_pts'], ex['pos_valid'], ex['neg_valid']
    G, R, K, _ = pos_pts.shape
    losses = {}
    num_pos = None
    if cfg['with_gt_loss']:
        gt_cls_prob = cls_prob(pos_cls[..., -1, :].reshape(G * R, -1), cfg)
        assert cfg['gt_loss_type'] == 'gt_refine'
        lab_rep = labels_all.unsqueeze(1).repeat(1, R).flatten()
        gt_valid = pos_valid[..., -1, :].reshape(G * R, -1)
        w_rep = gt_valid.float() * gt_weights.unsqueeze(1).repeat(1, R).flatten().reshape(-1, 1)
        onehot = torch.zeros_like(gt_cls_prob)
        onehot[torch.arange(len(onehot)), lab_rep] = 1
        num_pos = max((w_rep > 0).sum(), 1)
        gl = gfocal_loss(gt_cls_prob, onehot, w_rep, cfg['mil_eps'])
        losses['gt_loss'] = cfg['gt_loss_weight'] * (gl.sum() / num_pos)
    if cfg['with_mil_loss']:
        pol = cfg['refine_bag_policy']
        if pol == 'independent_with_gt_bag':
            rs = lambda t: t.reshape(G * R, K, -1)
            pw = gt_weights.unsqueeze(1).repeat(1, R).flatten()
            lab = labels_all.unsqueeze(1).repeat(1, R).flatten()
            c_, i_, v_ = rs(pos_cls), rs(pos_ins), rs(pos_valid)
        elif pol == 'merge_to_gt_bag':
            rs = lambda t: t.reshape(G, R * K, -1)
            pw, lab = gt_weights, labels_all
            c_, i_, v_ = rs(pos_cls), rs(pos_ins), rs(pos_valid)
        elif pol == 'only_refine_bag':
            si = 1 if R > 1 else 0
            rs = lambda t: t[:, si:].reshape(G, (R - si) * K, -1)
            pw, lab = gt_weights, labels_all
            c_, i_, v_ = rs(pos_cls), rs(pos_ins), rs(pos_valid)
        else:
            raise ValueError(pol)
        pos_w = v_.float() * pw.reshape(-1, 1, 1)
        # random_remove (ref:1119-1129,1213) only zeroes the unused stride column: no effect, skipped.
        pos_loss, acc, num_pos, bag_prob = mil_loss(cls_prob(c_, cfg), i_, lab, pos_w,
                                                    cfg['mil_loss_weight'], cfg['mil_eps'], cfg.get('binary_ins', False))
        losses['pos_loss'] = pos_loss
        losses['bag_acc'] = acc
    if cfg['with_neg']:
        neg_prob = cls_prob(neg_cls, cfg)
        nl = gfocal_loss(neg_prob, torch.zeros_like(neg_prob), neg_valid.float(), cfg['mil_eps'])
        losses['neg_loss'] = cfg['neg_loss_weight'] * (nl.sum() / num_pos)
    if return_all:
        return losses, dict(ex=ex, pos_cls=pos_cls, pos_ins=pos_ins, neg_cls=neg_cls,
                            bag_prob=bag_prob if cfg['with_mil_loss'] else None)
    return losses


# ----------------------------------------------------------------------------------------------
# PointRefiner (ref:665-895)
# ----------------------------------------------------------------------------------------------
def _group_indices(labels):
    groups = OrderedDict()
    for i, l in enumerate(labels.tolist()):
        groups.setdefault(l, []).append(i)
    return groups


def nearest_filter(bag_pts, gt_r_pts, gt_labels):
    """ref:711-743 class-wise: a bag point stays valid iff its nearest same-class GT(-refine) centre is
    its own one; classes with a single GT are not filtered. bag_pts (n,R,K,3), gt_r_pts (n,R,1,3)."""
    n, R, K, _ = bag_pts.shape
    valid = torch.ones((n, R * K), dtype=torch.bool)
    for l, idx in _group_indices(gt_labels).items():
        if len(idx) > 1:
            bp, gp = bag_pts[idx], gt_r_pts[idx]
            dist = torch.cdist(bp.flatten(0, -2)[..., :2], gp.flatten(0, -2)[..., :2], p=2)
            closest = dist.min(dim=1)[1].reshape(len(idx) * R, K)
            cur = torch.arange(len(closest)).reshape(-1, 1)
            valid[idx] = (closest == cur).reshape(len(idx), R * K)
    return valid


def refine_single(bag_cls_prob, bag_pts, bag_valid, gt_r_points, gt_labels, img_shape, cfg, not_refine=None):
    """ref:780-850.  bag_cls_prob (n,R,K,C), bag_pts (n,R,K,3), bag_valid (n,R,K,1) bool,
    gt_r_points (n,R,2), gt_labels (n,), img_shape (h,w,..).
    returns dict(refine_pts (n,2), refine_scores (n,), not_refine bool (n,), chosen bool (n,R*K),
                 merge_valid bool (n,R*K), plus the intermediate filter masks)."""
    n, R, K, C = bag_cls_prob.shape
    gt_cls_prob = bag_cls_prob[..., -1:, :]
    gt_r_pts = bag_pts[..., -1:, :]
    assert (gt_r_pts[:, :, 0, :2] == gt_r_points[:, :1]).all()
    gi = torch.arange(n)
    merge_valid = bag_valid.reshape(n, R * K).bool().clone()
    masks = {}
    if cfg['nearest_filter']:
        masks['nearest'] = nearest_filter(bag_pts, gt_r_pts, gt_labels)
        merge_valid &= masks['nearest']
    if cfg['classify_filter']:
        # ref:745-756
        masks['classify'] = (bag_cls_prob.max(dim=-1)[1] == gt_labels.reshape(n, 1, 1)).reshape(n, R * K)
        merge_valid &= masks['classify']
    p = bag_cls_prob[gi, ..., gt_labels].reshape(n, R * K)
    pg = gt_cls_prob[gi, 0, ..., gt_labels].reshape(n, 1)
    masks['thr'] = (p > cfg['merge_th']) & (p > pg * cfg['gt_alpha'])
    merge_valid &= masks['thr']
    h, w = img_shape[:2]
    bp = bag_pts.reshape(n, R * K, -1)
    x, y = bp[..., 0], bp[..., 1]
    masks['inside'] = (x < w) & (x >= 0) & (y < h) & (y >= 0)      # ref:773-778
    merge_valid &= masks['inside']
    p = p * merge_valid.float()
    wgt = p / (p.sum(dim=1, keepdim=True) + 1e-8)
    refine_pts = (bp[..., :2] * wgt.unsqueeze(-1)).sum(dim=1)
    scores = p.sum(dim=-1) / ((p > 0).float().sum(dim=-1) + 1e-8)
    cur = scores < cfg['refine_th']
    not_refine = cur if not_refine is None else (not_refine | cur)
    refine_pts[not_refine] = gt_r_points[:, 0][not_refine]
    if cfg['return_score_type'] == 'max':
        scores = p.max(dim=-1)[0]
        scores[scores == 0] = cfg['refine_th'] / 2
    out = dict(refine_pts=refine_pts, refine_scores=scores, not_refine=not_refine, chosen=wgt > 0,
               merge_valid=merge_valid)
    out.update({'mask_' + k: v for k, v in masks.items()})
    return out


def fill_list_to_tensor(alist, default_value=-1):
    """ref:55-61: ragged list of (len_i, 2) -> (n, max_len, 2) padded with -1."""
    max_l = max(len(l) for l in alist)
    data = torch.full((len(alist), max_l) + tuple(alist[0].shape[1:]), float(default_value))
    for i, l in enumerate(alist):
        data[i, :len(l)] = l
    return data


def cpr_get_bboxes(cls_feat, weights, gt_bboxes, gt_labels, gt_anns_id, img_metas, cfg, rescale=False,
                   return_all=False, out_geo=False):
    """CPRHead.get_bboxes (ref:1231-1283) -> [(det (n,6), labels (n,))] per image; with other_info.out_geo the rows grow
    by the flattened geometry [refined point, chosen bag points...] padded with -1 per image (ref:855-866, 1262-1273)."""
    gt_points = pseudo_bbox_to_center(gt_bboxes)
    gt_r_points = [p.reshape(len(l), -1, *p.shape[1:]) for p, l in zip(gt_points, gt_labels)]
    ex = extract(cls_feat, gt_r_points, gt_labels, img_metas, cfg)
    bag_prob = cls_prob(pts_outs(ex['pos_feats'], weights, 'cls_out', cfg.get('num_cls_fcs', 0)), cfg)
    results, inter = [], []
    s = 0
    for b, n in enumerate(ex['pos_len']):
        r = refine_single(bag_prob[s:s + n], ex['pos_pts'][s:s + n], ex['pos_valid'][s:s + n],
                          gt_r_points[b], gt_labels[b], img_metas[b]['img_shape'], cfg)
        s += n
        boxes = center_to_pseudo_bbox([r['refine_pts']])[0]
        if rescale:
            boxes = boxes / boxes.new_tensor(img_metas[b]['scale_factor'])
        det = torch.cat([boxes, r['refine_scores'].unsqueeze(-1),
                         gt_anns_id[b].unsqueeze(-1).type_as(boxes)], dim=-1)
        if out_geo:
            bp = ex['pos_pts'][s - n:s].reshape(n, -1, 3)
            geos = [torch.cat([rp[None, :2], bp[i][ch][:, :2]], dim=0)
                    for i, (rp, ch) in enumerate(zip(r['refine_pts'], r['chosen']))]            # ref:855-866
            if rescale:
                geos = [gg / gg.new_tensor(img_metas[b]['scale_factor'][:2]) for gg in geos]    # ref:1288-1291
            geo = fill_list_to_tensor(geos)
            det = torch.cat([det, geo.reshape(len(geo), -1)], dim=-1)
        results.append((det, gt_labels[b]))
        inter.append(r)
    if return_all:
        return results, dict(ex=ex, bag_prob=bag_prob, refine=inter)
    return results
