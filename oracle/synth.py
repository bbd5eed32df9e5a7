"""Seeded synthetic inputs for the CPR / P2P hot path.  Test/bench infrastructure.

Everything is generated on the CPU torch generator (bit-reproducible for a fixed torch version) so that the
oracle, the golden fixtures and the CUDA path all see identical bits.
"""
import math
import torch

# name -> (B, Hpad, Wpad, img_h, img_w, stride, n_points, radius, num_classes, channels)
CPR_CONFIGS = {
    # BASELINE.json configs[0]: CPR-lite 256x256, 32 points, bs=1 (plumbing)
    'lite': dict(B=1, pad_hw=(256, 256), img_hw=(250, 253), stride=8, n=32, radius=5, num_classes=80, C=256),
    # small multi-image case with points near the borders (parity fixture)
    'mid': dict(B=2, pad_hw=(192, 288), img_hw=(187, 281), stride=8, n=24, radius=8, num_classes=80, C=256),
    # BASELINE.json configs[1]: CPR R50-FPN 1333x800 (pad 800x1344), 500 points/img, bs=8  -- HEADLINE
    'headline': dict(B=8, pad_hw=(800, 1344), img_hw=(800, 1333), stride=8, n=500, radius=8, num_classes=80, C=256),
    # BASELINE.json configs[4] per-GPU shard: 2000 points/img, bs=4 per GPU
    'cpr2000': dict(B=4, pad_hw=(800, 1344), img_hw=(800, 1333), stride=8, n=2000, radius=8, num_classes=80, C=256),
}


def sample_points(n, w, h, gen, min_sep=4.0, border=0.0):
    """uniform points with pairwise distance >= min_sep (rejection), fp32 (n,2) as (x,y)."""
    pts = torch.empty(0, 2)
    tries = 0
    while len(pts) < n:
        cand = torch.rand(n * 2, 2, generator=gen) * torch.tensor([w - 2 * border, h - 2 * border]) + border
        for c in cand:
            if len(pts) == 0 or torch.cdist(c[None].double(), pts.double()).min() >= min_sep:
                pts = torch.cat([pts, c[None]])
                if len(pts) == n:
                    break
        tries += 1
        assert tries < 100
    return pts.float()


def cpr_weights(C, num_classes, gen, stacked_convs=4, scale=1.0, with_towers=True):
    """state_dict-shaped weights with the reference's parameter names (SURVEY §5 checkpoint row)."""
    w = {}
    if with_towers:
        for i in range(stacked_convs):
            w[f'cls_convs.{i}.conv.weight'] = torch.randn(C, C, 3, 3, generator=gen) * (1.4 / math.sqrt(C * 9))
            w[f'cls_convs.{i}.gn.weight'] = 1 + 0.1 * torch.randn(C, generator=gen)
            w[f'cls_convs.{i}.gn.bias'] = 0.1 * torch.randn(C, generator=gen)
    w['cls_out.weight'] = torch.randn(num_classes, C, generator=gen) * 0.01 * scale
    w['cls_out.bias'] = torch.full((num_classes,), -math.log(99.0))
    w['ins_out.weight'] = torch.randn(num_classes, C, generator=gen) * 0.01 * scale
    w['ins_out.bias'] = torch.zeros(num_classes)
    return w


def cpr_inputs(name='lite', seed=1234, trained_like=True, with_towers=False, **over):
    """returns dict(cfgd, x|cls_feat, weights, gt_bboxes, gt_labels, gt_anns_id, img_metas).

    `cls_feat` (B,C,H,W) plays the role of the tower output.  With trained_like=True, class evidence is
    planted around every GT (a bump along the cls_out row of its label, centred a few px off the
    annotated point) and cls_out is scaled so probabilities span (0,1): this exercises the refine filters.
    """
    d = dict(CPR_CONFIGS[name])
    d.update(over)
    gen = torch.Generator().manual_seed(seed)
    B, (ph, pw), (ih, iw), s, n, C, ncls = d['B'], d['pad_hw'], d['img_hw'], d['stride'], d['n'], d['C'], d['num_classes']
    H, W = ph // s, pw // s
    weights = cpr_weights(C, ncls, gen, scale=8.0 if trained_like else 1.0, with_towers=with_towers)
    feat = torch.relu(torch.randn(B, C, H, W, generator=gen))
    gt_bboxes, gt_labels, gt_anns_id, img_metas = [], [], [], []
    aid = 0
    for b in range(B):
        pts = sample_points(n, pw, ph, gen)
        labels = torch.randint(0, ncls, (n,), generator=gen)
        if trained_like:
            true_c = pts + (torch.rand(n, 2, generator=gen) - 0.5) * 3 * s      # object centre != coarse point
            yy, xx = torch.meshgrid(torch.arange(H).float() * s + s / 2, torch.arange(W).float() * s + s / 2, indexing='ij')
            wc = weights['cls_out.weight']
            planted = torch.rand(n, generator=gen) < 0.8       # ~20 % of the GTs get no evidence -> not_refine path
            for g in range(n):
                if not planted[g]:
                    continue
                bump = torch.exp(-((xx - true_c[g, 0]) ** 2 + (yy - true_c[g, 1]) ** 2) / (2 * (2.5 * s) ** 2))
                direction = wc[labels[g]] / (wc[labels[g]] ** 2).sum()
                feat[b] += 9.0 * direction[:, None, None] * bump[None]
        gt_bboxes.append(torch.cat([pts - 8, pts + 8], dim=1))
        gt_labels.append(labels)
        gt_anns_id.append(torch.arange(aid, aid + n))
        aid += n
        img_metas.append(dict(pad_shape=(ph, pw, 3), img_shape=(ih, iw, 3), scale_factor=[1.0, 1.0, 1.0, 1.0]))
    return dict(cfgd=d, cls_feat=feat.contiguous(), weights=weights, gt_bboxes=gt_bboxes, gt_labels=gt_labels,
                gt_anns_id=gt_anns_id, img_metas=img_metas)


P2P_CONFIGS = {
    'lite': dict(B=1, pad_hw=(128, 128), img_hw=(125, 126), stride=4, n=12, num_classes=80, C=256),
    'mid': dict(B=2, pad_hw=(160, 224), img_hw=(157, 219), stride=4, n=20, num_classes=80, C=256),
    # BASELINE.json configs[2] ("P2B-shaped"): 1333x800 at stride 8 -> 16800 proposals, bs=16
    'headline': dict(B=16, pad_hw=(800, 1344), img_hw=(800, 1333), stride=8, n=100, num_classes=80, C=256),
}


def p2p_inputs(name='lite', seed=4321, **over):
    """head OUTPUT maps (cls_out (B,C,H,W) logits ~ N(-3,1.5) made tie-free, pts_out (B,2,H,W)) + GTs."""
    d = dict(P2P_CONFIGS[name])
    d.update(over)
    gen = torch.Generator().manual_seed(seed)
    B, (ph, pw), (ih, iw), s, n, ncls = d['B'], d['pad_hw'], d['img_hw'], d['stride'], d['n'], d['num_classes']
    H, W = ph // s, pw // s
    cls_out = torch.randn(B, ncls, H, W, generator=gen) * 1.5 - 3.0
    cls_out += (torch.arange(cls_out.numel()).reshape(cls_out.shape) % 9973).float() * 1e-6   # break ties
    pts_out = torch.randn(B, 2, H, W, generator=gen) * 1.5
    gt_bboxes, gt_labels, img_metas = [], [], []
    for b in range(B):
        pts = sample_points(n, iw, ih, gen)
        gt_bboxes.append(torch.cat([pts - 8, pts + 8], dim=1))
        gt_labels.append(torch.randint(0, ncls, (n,), generator=gen))
        img_metas.append(dict(pad_shape=(ph, pw, 3), img_shape=(ih, iw, 3), scale_factor=[1.0, 1.0, 1.0, 1.0]))
    return dict(cfgd=d, cls_out=cls_out, pts_out=pts_out, gt_bboxes=gt_bboxes, gt_labels=gt_labels, img_metas=img_metas)


def p2p_aug_inputs(name='lite', seed=2468):
    """test-time-augmentation / cropped-tile case for P2PHead.aug_test_bboxes (p2p_head.py:487-572): 4 "augmentations" of one image,
    each a (cls_out, pts_out) pair of ONE image + its meta.  aug 1 re-uses aug 0's maps with a small tile offset (heavily overlapping
    boxes -> the second NMS has work to do), aug 2 is a horizontally flipped view at scale 1.5, aug 3 a vertically flipped tile."""
    base = p2p_inputs(name, seed, B=1)
    other = p2p_inputs(name, seed + 1, B=1)
    third = p2p_inputs(name, seed + 2, B=1)
    ih, iw = base['cfgd']['img_hw']
    ph, pw = base['cfgd']['pad_hw']

    def meta(scale, flip, direction, tile):
        m = dict(pad_shape=(ph, pw, 3), img_shape=(ih, iw, 3), scale_factor=[scale] * 4, flip=flip, flip_direction=direction)
        if tile is not None:
            m['tile_offset'] = tile
        return [m]
    outs = [(base['cls_out'], base['pts_out']), (base['cls_out'] + 0.01, base['pts_out']),
            (other['cls_out'], other['pts_out']), (third['cls_out'], third['pts_out'])]
    metas = [meta(1.0, False, 'horizontal', None), meta(1.0, False, 'horizontal', (3, 2)),
             meta(1.5, True, 'horizontal', None), meta(1.0, True, 'vertical', (40, 24))]
    return dict(cfgd=base['cfgd'], outs=outs, metas=metas)
