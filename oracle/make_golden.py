"""Pins the oracle against the REAL reference and writes tests/golden/*.npz.  (test infrastructure)

Run in the build container only (needs /root/reference):   python -m oracle.make_golden
  1. imports the unmodified reference mmdet package through oracle/_mmcv_stub.py,
  2. builds CPRHead / P2PHead from the reference's own config dicts
     (configs2/_base_/models/cpr/coarse_point_refine_r50_fpns4_1x.py:26-69,
      configs2/COCO/p2p/p2p_r50_fpns4_1x_fl_sl1_coco.py:85-127),
  3. runs reference and oracle on the same seeded inputs (oracle/synth.py) and ASSERTS equality
     (bit-exact for bool/int outputs; exact-or-1e-6 for floats: both call the same ATen CPU kernels),
  4. stores the reference's outputs as golden vectors.
"""
import os
import sys
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import cpr as ocpr, p2p as op2p, synth  # noqa: E402
from oracle._mmcv_stub import load_reference, CfgDict  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')


def ref_cpr_cfg(d):
    r = d['radius']
    return dict(
        type='CPRHead', norm_cfg=dict(type='GN', num_groups=32, requires_grad=True),
        num_classes=d['num_classes'], in_channels=d['C'], feat_channels=d['C'], stacked_convs=4, num_cls_fcs=0,
        strides=[d['stride']],
        loss_mil=dict(type='MILLoss', binary_ins=False, loss_weight=0.25), loss_type=0,
        loss_cfg=dict(with_neg=True, neg_loss_weight=0.75, refine_bag_policy='only_refine_bag',
                      random_remove_rate=0.4, with_gt_loss=True, gt_loss_weight=0.125, with_mil_loss=True),
        normal_cfg=dict(prob_cls_type='sigmoid', out_bg_cls=False),
        train_pts_extractor=dict(pos_generator=dict(type='CirclePtFeatGenerator', radius=r),
                                 neg_generator=dict(type='OutCirclePtFeatGenerator', radius=r, class_wise=True)),
        refine_pts_extractor=dict(pos_generator=dict(type='CirclePtFeatGenerator', radius=r),
                                  neg_generator=dict(type='OutCirclePtFeatGenerator', radius=r, keep_wh=True,
                                                     class_wise=True)),
        point_refiner=dict(merge_th=0.1, refine_th=0.1, classify_filter=True, nearest_filter=True),
        train_cfg=None, test_cfg=CfgDict(nms_pre=1000, min_bbox_size=0, score_thr=0.05,
                                         nms=dict(type='nms', iou_threshold=0.5), max_per_img=100))


def eq(a, b, what, exact=True, tol=1e-6):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.dtype in (torch.bool, torch.int64, torch.int32, torch.uint8):
        assert torch.equal(a, b), f'{what}: integer mismatch {(a != b).sum().item()}'
        return 0.0
    d = (a.double() - b.double()).abs().max().item() if a.numel() else 0.0
    if exact:
        assert d == 0.0, f'{what}: float mismatch {d}'
    else:
        assert d <= tol * max(1.0, b.abs().max().item()), f'{what}: float mismatch {d}'
    return d


def sub(t, step=13):
    """strided sample of a big float tensor + float64 checksum, to keep fixtures small."""
    f = t.detach().flatten()
    return f[::step].numpy().copy(), np.float64(f.double().sum().item()), np.float64(f.double().abs().sum().item())


def golden_cpr(HEADS, name, seed, with_towers=False, grid_radius=None):
    inp = synth.cpr_inputs(name, seed, trained_like=True, with_towers=with_towers)
    d = inp['cfgd']
    cfg = ocpr.default_cfg(num_classes=d['num_classes'], in_channels=d['C'], feat_channels=d['C'], stride=d['stride'],
                           pos_radius=d['radius'], neg_radius=d['radius'])
    rcfg = ref_cpr_cfg(d)
    if grid_radius is not None:
        # a8: grid-cell bags (GridCirclesPtFeatGenerator, cpr_head.py:413-444) instead of ring bags, train + refine
        for ex_name in ('train_pts_extractor', 'refine_pts_extractor'):
            rcfg[ex_name]['pos_generator'] = dict(type='GridCirclesPtFeatGenerator', radius=grid_radius)
        cfg.update(pos_generator='grid_circles', pos_radius=grid_radius)
    head = HEADS.build(rcfg)
    sd = head.state_dict()
    w = dict(inp['weights'])
    if not with_towers:
        for k in sd:
            if k.startswith('cls_convs'):
                w[k] = sd[k]
    missing = head.load_state_dict(w, strict=True)
    head.eval()
    feat = inp['cls_feat']
    gtb, gtl, metas, aid = inp['gt_bboxes'], inp['gt_labels'], inp['img_metas'], inp['gt_anns_id']
    out = {}
    # ---- towers
    if with_towers:
        x = feat
        with torch.no_grad():
            ref_cls_feat = head((x,))[0][0]
            ora = ocpr.tower_forward(x, w, cfg)
        eq(ora, ref_cls_feat, 'tower', exact=False, tol=1e-5)
        out['tower_sub'], out['tower_sum'], out['tower_abs'] = sub(ref_cls_feat, 97)
        feat = ref_cls_feat
    # ---- extraction (train extractor)
    gt_points = head.pseudo_bbox_to_center(gtb)
    gt_r = [p.reshape(len(l), -1, *p.shape[1:]) for p, l in zip(gt_points, gtl)]
    with torch.no_grad():
        pos_data, neg_data = head.train_pts_extractor([feat], [feat], gt_r, gtl, metas, None, True)
        ex = ocpr.extract(feat, gt_r, gtl, metas, cfg)
        eq(ex['pos_pts'], pos_data.pts[0], 'pos_pts')
        eq(ex['pos_valid'], pos_data.valid[0], 'pos_valid')
        eq(ex['neg_valid'], neg_data.valid[0], 'neg_valid')
        eq(ex['neg_pts'], neg_data.pts[0], 'neg_pts')
        eq(ex['pos_feats'], pos_data.cls_feats[0], 'pos_feats')
        eq(ex['neg_feats'], neg_data.cls_feats[0], 'neg_feats')
        ref_pos_cls, ref_pos_ins = head.get_pts_outs(pos_data.cls_feats, pos_data.ins_feats)
        eq(ocpr.pts_outs(ex['pos_feats'], w, 'cls_out'), ref_pos_cls[0], 'pos_cls')
        eq(ocpr.pts_outs(ex['pos_feats'], w, 'ins_out'), ref_pos_ins[0], 'pos_ins')
    out['pos_valid'] = pos_data.valid[0].numpy()
    if grid_radius is not None:
        gen = head.train_pts_extractor.pos_generator
        ch_all = []
        for b in range(len(metas)):
            H, W = feat.shape[2:]
            gpts, _ = gen.anchor_points(H, W, *metas[b]['pad_shape'][:2], d['stride'], feat.device)
            ch, _ = gen.get_chosen_neighbours(gpts, gt_r[b], d['stride'])
            _, _, _, och = ocpr.grid_circles_bag(feat[b:b + 1], gt_r[b], *metas[b]['pad_shape'][:2], d['stride'], grid_radius,
                                                 keep_feats=False)
            eq(och, ch, f'chosens[{b}]')
            ch_all.append(ch.reshape(len(ch), -1))
        ch_all = torch.cat(ch_all)
        out['chosens'] = np.packbits(ch_all.numpy(), axis=None)
        out['chosens_shape'] = np.array(ch_all.shape)
        out['grid_radius'] = np.int64(grid_radius)
    out['neg_valid'] = np.packbits(neg_data.valid[0].numpy(), axis=None)
    out['neg_valid_shape'] = np.array(neg_data.valid[0].shape)
    out['pos_pts'] = pos_data.pts[0].numpy()
    out['pos_feats_sub'], out['pos_feats_sum'], out['pos_feats_abs'] = sub(pos_data.cls_feats[0], 1009)
    out['pos_cls_sub'], out['pos_cls_sum'], out['pos_cls_abs'] = sub(ref_pos_cls[0], 101)
    out['pos_ins_sub'], out['pos_ins_sum'], out['pos_ins_abs'] = sub(ref_pos_ins[0], 101)
    # ---- loss (+ gradients w.r.t. feature map and FC weights)
    feat_g = feat.clone().requires_grad_(True)
    head.zero_grad()
    ref_losses = head.loss([feat_g], [feat_g], gtb, gtl, metas)
    total = sum(v for k, v in ref_losses.items() if 'loss' in k)
    total.backward()
    feat_o = feat.clone().requires_grad_(True)
    wo = {k: v.clone().requires_grad_(k.startswith(('cls_out', 'ins_out'))) for k, v in w.items()}
    ora_losses, ora_all = ocpr.cpr_loss(feat_o, wo, gtb, gtl, metas, cfg, return_all=True)
    sum(v for k, v in ora_losses.items() if 'loss' in k).backward()
    for k in ref_losses:
        eq(ora_losses[k].detach().reshape(-1), ref_losses[k].detach().reshape(-1), 'loss ' + k, exact=False, tol=1e-6)
        out['loss_' + k] = ref_losses[k].detach().reshape(-1).numpy()
    eq(feat_o.grad, feat_g.grad, 'dfeat', exact=False, tol=1e-5)
    eq(wo['cls_out.weight'].grad, head.cls_out.weight.grad, 'dWcls', exact=False, tol=1e-5)
    eq(wo['ins_out.weight'].grad, head.ins_out.weight.grad, 'dWins', exact=False, tol=1e-5)
    out['grad_feat_sub'], out['grad_feat_sum'], out['grad_feat_abs'] = sub(feat_g.grad, 211)
    out['grad_cls_w'] = head.cls_out.weight.grad.numpy()
    out['grad_cls_b'] = head.cls_out.bias.grad.numpy()
    out['grad_ins_w'] = head.ins_out.weight.grad.numpy()
    out['grad_ins_b'] = head.ins_out.bias.grad.numpy()
    out['mil_bag_prob'] = ora_all['bag_prob'].detach().numpy()
    # ---- refine (get_bboxes); also dig the PointRefiner intermediates out of the reference
    with torch.no_grad():
        ref_res = head.get_bboxes([feat], [feat], metas, gt_bboxes=gtb, gt_labels=gtl, gt_anns_id=aid)
        bag_data, grid_data = head.refine_pts_extractor([feat], [feat], gt_r, gtl, metas, None, True)
        bag_data.cls_outs, bag_data.ins_outs = head.get_pts_outs(bag_data.cls_feats, bag_data.ins_feats)
        grid_data.cls_outs = head.get_pts_outs(grid_data.cls_feats)
        bag_data.cls_prob = [head.get_cls_prob(o) for o in bag_data.cls_outs]
        grid_data.cls_prob = [head.get_cls_prob(o) for o in grid_data.cls_outs]
        r_pts, r_scores, r_not, r_geos = head.point_refiner(bag_data, grid_data, gt_r, gtl, metas, None, None)
        ora_res, ora_all = ocpr.cpr_get_bboxes(feat, w, gtb, gtl, aid, metas, cfg, return_all=True)
    for b in range(len(metas)):
        eq(ora_res[b][0], ref_res[b][0], f'det[{b}]')
        eq(ora_all['refine'][b]['not_refine'], r_not[b], f'not_refine[{b}]')
        eq(ora_all['refine'][b]['refine_scores'], r_scores[b], f'scores[{b}]')
        chosen_n = torch.tensor([len(g) - 1 for g in r_geos[b]])
        eq(ora_all['refine'][b]['chosen'].sum(1), chosen_n, f'chosen count[{b}]')
    out['det'] = torch.cat([r[0] for r in ref_res]).numpy()
    out['not_refine'] = torch.cat(r_not).numpy()
    out['chosen'] = torch.cat([r['chosen'] for r in ora_all['refine']]).numpy()
    out['merge_valid'] = torch.cat([r['merge_valid'] for r in ora_all['refine']]).numpy()
    out['mask_nearest'] = torch.cat([r['mask_nearest'] for r in ora_all['refine']]).numpy()
    out['mask_classify'] = torch.cat([r['mask_classify'] for r in ora_all['refine']]).numpy()
    out['frac_not_refine'] = np.float64(out['not_refine'].mean())
    out['seed'] = np.int64(seed)
    path = os.path.join(GOLD, f'cpr_{name}{"_tower" if with_towers else ""}{"_grid" if grid_radius is not None else ""}.npz')
    np.savez_compressed(path, **out)
    print(f'[golden] {path}: {os.path.getsize(path) / 1024:.0f} KiB; not_refine frac {out["frac_not_refine"]:.3f}; '
          f'mean chosen/bag {out["chosen"].sum(1).mean():.1f}; losses '
          + ' '.join(f'{k}={float(v.reshape(-1)[0]):.5f}' for k, v in ref_losses.items()))


CPR_VARIANTS = {
    # name -> (reference ctor overrides, oracle cfg overrides)
    'softmax': (dict(normal_cfg=dict(prob_cls_type='softmax', out_bg_cls=False)), dict(prob_cls_type='softmax')),
    'normed_sigmoid': (dict(normal_cfg=dict(prob_cls_type='normed_sigmoid', out_bg_cls=False, normed_sigmoid_p=2)),
                       dict(prob_cls_type='normed_sigmoid', normed_sigmoid_p=2)),
    'fcs_binary': (dict(num_cls_fcs=2, fc_out_channels=64, loss_mil=dict(type='MILLoss', binary_ins=True, loss_weight=0.25)),
                   dict(num_cls_fcs=2, binary_ins=True)),
}


def variant_weights(inp, variant, seed):
    """extra parameters of a variant head (FC stack, doubled instance classifier) under the reference's names, seeded."""
    g = torch.Generator().manual_seed(seed + 17)
    w = dict(inp['weights'])
    C = inp['cfgd']['C']
    if variant == 'fcs_binary':
        chn = C
        for i in range(2):
            w[f'cls_fcs.{i}.weight'] = torch.randn(64, chn, generator=g) * (1.0 / chn ** 0.5)
            w[f'cls_fcs.{i}.bias'] = torch.randn(64, generator=g) * 0.1
            chn = 64
        n = inp['cfgd']['num_classes']
        w['cls_out.weight'] = torch.randn(n, 64, generator=g) * 0.3
        w['ins_out.weight'] = torch.randn(2 * n, 64, generator=g) * 0.3
        w['ins_out.bias'] = torch.zeros(2 * n)
    return w


def golden_cpr_variant(HEADS, name, seed, variant):
    """non-default CPRHead variants (VERDICT r1 'missing' #1): the REAL reference head built with the variant's ctor kwargs; loss (+
    gt_weights), gradients and get_bboxes compared with the oracle restatement, stored as golden vectors."""
    inp = synth.cpr_inputs(name, seed, trained_like=True)
    d = inp['cfgd']
    ref_over, ora_over = CPR_VARIANTS[variant]
    rcfg = ref_cpr_cfg(d)
    rcfg.update(ref_over)
    head = HEADS.build(rcfg)
    w = variant_weights(inp, variant, seed)
    sd = head.state_dict()
    for k in sd:
        if k.startswith('cls_convs'):
            w[k] = sd[k]
    head.load_state_dict(w, strict=True)
    head.eval()
    cfg = ocpr.default_cfg(num_classes=d['num_classes'], in_channels=d['C'], feat_channels=d['C'], stride=d['stride'], pos_radius=d['radius'],
                           neg_radius=d['radius'], **ora_over)
    gtb, gtl, metas, aid = inp['gt_bboxes'], inp['gt_labels'], inp['img_metas'], inp['gt_anns_id']
    g = torch.Generator().manual_seed(seed + 5)
    gtw = [torch.rand(len(l), generator=g) * 0.5 + 0.5 for l in gtl]
    gtw[0][0] = 0.0                                               # a bag whose weight is zero
    out = {}
    f_ref = inp['cls_feat'].clone().requires_grad_(True)
    rl = head.loss([f_ref], [f_ref], gtb, gtl, metas, gt_weights=gtw)
    sum(v for k, v in rl.items() if 'loss' in k).backward()
    f_o = inp['cls_feat'].clone().requires_grad_(True)
    wo = {k: v.clone().requires_grad_(True) for k, v in w.items() if not k.startswith('cls_convs')}
    ol = ocpr.cpr_loss(f_o, wo, gtb, gtl, metas, cfg, gt_weights=gtw)
    sum(v for k, v in ol.items() if 'loss' in k).backward()
    for k in ('gt_loss', 'pos_loss', 'neg_loss', 'bag_acc'):
        eq(ol[k].detach().reshape(-1), rl[k].detach().reshape(-1), f'{variant} {k}', exact=False, tol=1e-6)
        out['loss_' + k] = rl[k].detach().reshape(-1).numpy()
    eq(f_o.grad, f_ref.grad, f'{variant} dfeat', exact=False, tol=1e-6)
    eq(wo['cls_out.weight'].grad, head.cls_out.weight.grad, f'{variant} dWcls', exact=False, tol=1e-6)
    out['grad_feat_sub'], out['grad_feat_sum'], out['grad_feat_abs'] = sub(f_ref.grad, 211)
    out['grad_cls_w'] = head.cls_out.weight.grad.numpy()
    with torch.no_grad():
        rres = head.get_bboxes([inp['cls_feat']], [inp['cls_feat']], metas, gt_bboxes=gtb, gt_labels=gtl, gt_anns_id=aid)
        ores, oall = ocpr.cpr_get_bboxes(inp['cls_feat'], w, gtb, gtl, aid, metas, cfg, return_all=True)
    for b in range(len(rres)):
        eq(ores[b][0], rres[b][0], f'{variant} det[{b}]')
    out['det'] = torch.cat([r[0] for r in rres]).numpy()
    out['not_refine'] = torch.cat([r['not_refine'] for r in oall['refine']]).numpy()
    out['chosen'] = np.packbits(torch.cat([r['chosen'] for r in oall['refine']]).numpy(), axis=None)
    path = os.path.join(GOLD, f'cpr_{name}_{variant}.npz')
    np.savez_compressed(path, **out)
    print(f'[golden] {path}: losses ' + ' '.join(f'{k}={float(v.reshape(-1)[0]):.5f}' for k, v in rl.items()) +
          f'; not_refine {float(out["not_refine"].mean()):.2f}')


def ref_p2p_cfg(d):
    return dict(
        type='P2PHead', norm_cfg=dict(type='GN', num_groups=32, requires_grad=True),
        num_classes=d['num_classes'], in_channels=d['C'], feat_channels=d['C'], stacked_convs=4, strides=[d['stride']],
        point_anchor=[(0., 0.)],
        loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0),
        loss_reg=dict(type='SmoothL1Loss', beta=1.0 / 9.0, loss_weight=0.5), pts_gamma=1, reg_norm=1,
        train_cfg=CfgDict(neg_weight=1.0,
                          assigner=dict(type='HungarianAssignerV2', cls_costs=dict(type='FocalLossCost', weight=2.0),
                                        reg_costs=dict(type='DisCostV2', weight=0.1, norm_with_img_wh=False), topk_k=5),
                          sampler=dict(type='PseudoSampler')),
        test_cfg=CfgDict(nms_pre=1000, min_bbox_size=0, score_thr=0.05, pseudo_wh=(32, 32),
                         nms=dict(type='nms', iou_threshold=0.01), max_per_img=100))


def golden_p2p(HEADS, name, seed, nms_iou=0.01):
    import mmdet.models.point.dense_heads.p2p_head as ref_mod
    ref_mod.TestP2PHead.test_assign = staticmethod(lambda *a, **k: None)   # debug visualiser (needs huicv)
    inp = synth.p2p_inputs(name, seed)
    d = inp['cfgd']
    rcfg = ref_p2p_cfg(d)
    rcfg['test_cfg']['nms'] = dict(type='nms', iou_threshold=nms_iou)
    head = HEADS.build(rcfg)
    cfg = op2p.default_cfg(num_classes=d['num_classes'], stride=d['stride'], nms_iou=nms_iou)
    cls_out, pts_out = inp['cls_out'], inp['pts_out']
    gtb, gtl, metas = inp['gt_bboxes'], inp['gt_labels'], inp['img_metas']
    out = {}
    with torch.no_grad():
        ra, rp, rv, rc = head.get_pred_points([cls_out], [pts_out], metas)
        oa, opd, ov, oc = op2p.pred_points(cls_out, pts_out, metas, cfg)
        eq(oa, ra, 'anchor'); eq(opd, rp, 'pred'); eq(ov, rv, 'valid'); eq(oc, rc, 'cls')
        # targets
        gt_points = head.pseudo_bbox_to_center(gtb)
        rl, rlw, rgp, rpw = head.get_targets(rp[..., :2], rv, rc, gt_points, gtl, metas, None)
        tg = [op2p.target_single(opd[b][..., :2], ov[b], oc[b], gt_points[b], gtl[b], metas[b]['img_shape'], cfg)
              for b in range(len(metas))]
        for b in range(len(metas)):
            eq(tg[b][0], rl[b], 'labels'); eq(tg[b][1], rlw[b], 'lw'); eq(tg[b][2], rgp[b], 'gpts'); eq(tg[b][3], rpw[b], 'pw')
        out['gt_inds'] = torch.stack([t[4] for t in tg]).numpy().astype(np.int32)
        # cost matrix fixture (first image)
        cm = op2p.cost_matrix(opd[0][ov[0]][..., :2], oc[0][ov[0]], gt_points[0], gtl[0], metas[0]['img_shape'], cfg)
        out['cost_sub'], out['cost_sum'], out['cost_abs'] = sub(cm, 37)
    co = cls_out.clone().requires_grad_(True)
    po = pts_out.clone().requires_grad_(True)
    rloss = head.loss([co], [po], gtb, gtl, metas, gt_bboxes_ignore=None) if False else None
    co_r = cls_out.clone().requires_grad_(True); po_r = pts_out.clone().requires_grad_(True)
    ign = [torch.zeros(0, 4) for _ in metas]
    rloss = head.loss([co_r], [po_r], gtb, gtl, metas, gt_bboxes_ignore=ign)
    (sum(rloss['loss_cls']) + sum(rloss['loss_pts'])).backward()
    oloss = op2p.p2p_loss(co, po, gtb, gtl, metas, cfg)
    (sum(oloss['loss_cls']) + sum(oloss['loss_pts'])).backward()
    for k in ('loss_cls', 'loss_pts'):
        eq(torch.stack(oloss[k]).detach(), torch.stack(rloss[k]).detach(), k, exact=False, tol=1e-6)
        out[k] = torch.stack(rloss[k]).detach().numpy()
    eq(co.grad, co_r.grad, 'dcls', exact=False, tol=1e-6)
    eq(po.grad, po_r.grad, 'dpts', exact=False, tol=1e-6)
    out['grad_cls_sub'], out['grad_cls_sum'], out['grad_cls_abs'] = sub(co_r.grad, 211)
    out['grad_pts_sub'], out['grad_pts_sum'], out['grad_pts_abs'] = sub(po_r.grad, 7)
    with torch.no_grad():
        rres = head.get_bboxes([cls_out], [pts_out], metas)
        ores = op2p.p2p_get_bboxes(cls_out, pts_out, metas, cfg)
        dets, labs, keeps, cands, topks = [], [], [], [], []
        for b in range(len(metas)):
            eq(ores[b][0], rres[b][0], f'p2p det[{b}]')
            eq(ores[b][1], rres[b][1], f'p2p labels[{b}]')
            _, _, al = op2p.get_bboxes_single(opd[b][..., :2], oc[b], metas[b]['img_shape'], metas[b]['scale_factor'],
                                              cfg, return_all=True)
            dets.append(rres[b][0]); labs.append(rres[b][1]); keeps.append(al['keep']); cands.append(al['cand_inds'])
            topks.append(al['topk_inds'] if al['topk_inds'] is not None else torch.zeros(0, dtype=torch.long))
    out['det_len'] = np.array([len(x) for x in dets])
    out['det'] = torch.cat(dets).numpy()
    out['det_labels'] = torch.cat(labs).numpy()
    out['keep'] = torch.cat(keeps).numpy()
    out['cand_len'] = np.array([len(x) for x in cands])
    out['cand'] = torch.cat(cands).numpy().astype(np.int32)
    out['topk'] = torch.cat(topks).numpy().astype(np.int32)
    out['seed'] = np.int64(seed)
    path = os.path.join(GOLD, f'p2p_{name}_iou{nms_iou}.npz')
    np.savez_compressed(path, **out)
    print(f'[golden] {path}: {os.path.getsize(path) / 1024:.0f} KiB; dets/img {out["det_len"].tolist()} '
          f'cands/img {out["cand_len"].tolist()} pos {int((out["gt_inds"] > 0).sum())}')


def golden_p2p_aug(HEADS, name, seed, nms_iou=0.5):
    """P2PHead.aug_test_bboxes (p2p_head.py:487-572) of the REAL reference; `forward` is replaced by a table lookup of prepared head
    outputs (the towers are pinned elsewhere), everything after it — per-aug get_bboxes + NMS, score scatter, bbox_mapping_back with
    flip / scale / tile_offset, the second multiclass_nms, the un-rescale — is the reference's own code."""
    inp = synth.p2p_aug_inputs(name, seed)
    d = inp['cfgd']
    rcfg = ref_p2p_cfg(d)
    rcfg['test_cfg']['nms'] = dict(type='nms', iou_threshold=nms_iou)
    head = HEADS.build(rcfg)
    table = {id(o[0]): o for o in inp['outs']}
    head.forward = lambda x: ([table[id(x)][0]], [table[id(x)][1]])
    cfg = op2p.default_cfg(num_classes=d['num_classes'], stride=d['stride'], nms_iou=nms_iou)
    out = {}
    for rescale in (False, True):
        with torch.no_grad():
            rres = head.aug_test_bboxes([o[0] for o in inp['outs']], inp['metas'], rescale=rescale)
            ores, aux = op2p.aug_test_bboxes(inp['outs'], inp['metas'], cfg, rescale=rescale)
        assert len(rres) == 1
        eq(ores[0][0], rres[0][0], f'aug det (rescale={rescale})')
        eq(ores[0][1], rres[0][1], f'aug labels (rescale={rescale})')
        out[f'det_rescale{int(rescale)}'] = rres[0][0].numpy()
        out[f'labels_rescale{int(rescale)}'] = rres[0][1].numpy()
    out['keep'] = aux['keep'].numpy()
    out['n_merged'] = np.int64(len(aux['merged_boxes']))
    path = os.path.join(GOLD, f'p2p_aug_{name}.npz')
    np.savez_compressed(path, **out)
    print(f'[golden] {path}: merged {int(out["n_merged"])} boxes -> {len(out["keep"])} kept')


def golden_multiclass_nms_options():
    """core/post_processing/bbox_nms.py:7-94 of the REAL reference (batched_nms from the mmcv stub = oracle/p2p.py's restatement of the
    third-party op) with the options the heads never use: score_factors, class-specific boxes (n, C*4), class_agnostic, max_num=-1."""
    from mmdet.core.post_processing.bbox_nms import multiclass_nms as ref_mnms
    g = torch.Generator().manual_seed(8642)
    n, C = 160, 6
    ctr = torch.rand(n, 2, generator=g) * 200
    wh = torch.rand(n, 2, generator=g) * 40 + 8
    boxes = torch.cat([ctr - wh / 2, ctr + wh / 2], 1)
    jit = torch.randn(n, C, 4, generator=g) * 3
    boxes_cs = (boxes[:, None] + jit).reshape(n, C * 4)
    scores = torch.rand(n, C + 1, generator=g) ** 3
    factors = torch.rand(n, generator=g)
    out = dict(boxes=boxes.numpy(), boxes_cs=boxes_cs.numpy(), scores=scores.numpy(), factors=factors.numpy())
    cases = dict(factors=dict(b=boxes, sf=factors, cfg=dict(type='nms', iou_threshold=0.5), max_num=50),
                 class_specific=dict(b=boxes_cs, sf=None, cfg=dict(type='nms', iou_threshold=0.4), max_num=80),
                 agnostic=dict(b=boxes, sf=None, cfg=dict(type='nms', iou_threshold=0.5, class_agnostic=True), max_num=60),
                 unlimited=dict(b=boxes, sf=factors, cfg=dict(type='nms', iou_threshold=0.3), max_num=-1))
    for k, c in cases.items():
        rd, rl, rk = ref_mnms(c['b'], scores, 0.05, dict(c['cfg']), c['max_num'], score_factors=c['sf'], return_inds=True)
        od, ol, ok, _ = op2p.multiclass_nms(c['b'], scores, 0.05, c['cfg']['iou_threshold'], c['max_num'], nms_cfg=c['cfg'], score_factors=c['sf'])
        eq(od, rd, f'multiclass_nms[{k}] dets'); eq(ol, rl, f'multiclass_nms[{k}] labels'); eq(ok, rk, f'multiclass_nms[{k}] keep')
        out[f'{k}_dets'] = rd.numpy(); out[f'{k}_labels'] = rl.numpy(); out[f'{k}_keep'] = rk.numpy()
    path = os.path.join(GOLD, 'multiclass_nms_options.npz')
    np.savez_compressed(path, **out)
    print(f'[golden] {path}: ' + ', '.join(f'{k} {len(out[k + "_keep"])} kept' for k in cases))


def golden_point_assigner():
    """the reference's own KATs: tests/test_utils/test_assigner.py:155-194."""
    pts = torch.FloatTensor([[0, 0, 1], [10, 10, 1], [5, 5, 1], [32, 32, 1]])
    gts = torch.FloatTensor([[0, 0, 10, 9], [0, 10, 10, 19]])
    from mmdet.core.bbox.assigners import PointAssigner
    r = PointAssigner().assign(pts, gts).gt_inds
    assert r.tolist() == [1, 2, 1, 0]
    assert op2p.point_assigner(pts, gts).tolist() == [1, 2, 1, 0]
    assert op2p.point_assigner(pts, torch.zeros(0, 4)).tolist() == [0, 0, 0, 0]
    assert len(op2p.point_assigner(torch.zeros(0, 3), torch.zeros(0, 4))) == 0
    g = torch.Generator().manual_seed(7)
    lv = torch.tensor([8., 16., 32.])
    pts = torch.cat([torch.rand(300, 2, generator=g) * 200, lv[torch.randint(0, 3, (300,), generator=g)][:, None]], 1)
    c = torch.rand(12, 2, generator=g) * 180 + 10
    wh = torch.rand(12, 2, generator=g) * 120 + 6
    gts = torch.cat([c - wh / 2, c + wh / 2], 1)
    r = PointAssigner(scale=4, pos_num=3).assign(pts, gts).gt_inds
    assert torch.equal(op2p.point_assigner(pts, gts), r)
    np.savez_compressed(os.path.join(GOLD, 'point_assigner.npz'), points=pts.numpy(), gts=gts.numpy(), gt_inds=r.numpy())
    print('[golden] point_assigner ok; assigned', int((r > 0).sum()))


def golden_result_json(HEADS):
    """CPR output -> annotation hand-off: reference get_bboxes (out_geo) -> reference bbox2result ->
    reference CocoDataset._det2json, stored as json; pins pointtinybenchmark_b200/results.py."""
    import json
    import types
    from mmdet.core.bbox.transforms import bbox2result
    from mmdet.datasets.coco import CocoDataset
    inp = synth.cpr_inputs('lite', 1234, trained_like=True)
    d = inp['cfgd']
    rcfg = ref_cpr_cfg(d)
    rcfg['other_info'] = dict(out_geo=True)
    head = HEADS.build(rcfg)
    w = dict(inp['weights'])
    for k, v in head.state_dict().items():
        if k.startswith('cls_convs'):
            w[k] = v
    head.load_state_dict(w, strict=True)
    head.eval()
    feat = inp['cls_feat']
    metas = [dict(m, scale_factor=[1.5, 1.25, 1.5, 1.25]) for m in inp['img_metas']]
    with torch.no_grad():
        res = head.get_bboxes([feat], [feat], metas, rescale=True, gt_bboxes=inp['gt_bboxes'], gt_labels=inp['gt_labels'],
                              gt_anns_id=inp['gt_anns_id'])
    cfg = ocpr.default_cfg(num_classes=d['num_classes'], in_channels=d['C'], feat_channels=d['C'], stride=d['stride'],
                           pos_radius=d['radius'], neg_radius=d['radius'])
    ora = ocpr.cpr_get_bboxes(feat, w, inp['gt_bboxes'], inp['gt_labels'], inp['gt_anns_id'], metas, cfg, rescale=True,
                              out_geo=True)
    for a, b in zip(res, ora):
        eq(b[0], a[0], 'det with geo')
    per_class = [bbox2result(dt, lb, d['num_classes']) for dt, lb in res]
    img_ids = [1000 + i for i in range(len(res))]
    cat_ids = [10 * (c + 1) for c in range(d['num_classes'])]
    class _Len(types.SimpleNamespace):
        def __len__(self):
            return len(self.img_ids)
    fake = _Len(img_ids=img_ids, cat_ids=cat_ids)
    fake.xyxy2xywh = lambda b: CocoDataset.xyxy2xywh(fake, b)
    js = CocoDataset._det2json(fake, per_class)
    path = os.path.join(GOLD, 'cpr_result_json.json')
    # mmcv.dump's json handler serialises numpy scalars with .item() (mmcv/fileio/handlers/json_handler.py set_default)
    json.dump(dict(img_ids=img_ids, cat_ids=cat_ids, scale_factor=[1.5, 1.25, 1.5, 1.25], results=js), open(path, 'w'),
              default=lambda o: o.item())
    print(f'[golden] {path}: {len(js)} result rows, {os.path.getsize(path) / 1024:.0f} KiB; geo lens',
          sorted({len(r["geo"]) for r in js})[:6])


def golden_max_iou():
    """dense-anchor assignment: the real MaxIoUAssigner / bbox_overlaps on seeded anchor sets."""
    from mmdet.core.bbox.assigners import MaxIoUAssigner
    from mmdet.core.bbox.iou_calculators import bbox_overlaps as ref_ov
    from oracle import anchors as oa
    MAX_IOU_CFGS = oa.MAX_IOU_CFGS
    out = {}
    for seed in (1, 2):
        a, g, l, ign = oa.synth_anchor_case(seed)
        eq(oa.bbox_overlaps(g, a), ref_ov(g, a), 'iou matrix')
        eq(oa.bbox_overlaps(a, ign, 'iof'), ref_ov(a, ign, mode='iof'), 'iof matrix')
        out[f's{seed}_iou_sub'], out[f's{seed}_iou_sum'], _ = sub(ref_ov(g, a), 7)
        for ci, kw in enumerate(MAX_IOU_CFGS):
            r = MaxIoUAssigner(**kw).assign(a, g, gt_bboxes_ignore=ign, gt_labels=l)
            gi, mo, lb = oa.max_iou_assign(a, g, l, ign, **kw)
            eq(gi, r.gt_inds, 'gt_inds'); eq(mo, r.max_overlaps, 'max_overlaps'); eq(lb, r.labels, 'labels')
            out[f's{seed}_c{ci}_gt_inds'] = r.gt_inds.numpy().astype(np.int16)
            out[f's{seed}_c{ci}_labels'] = r.labels.numpy().astype(np.int16)
            out[f's{seed}_c{ci}_max_overlaps'] = r.max_overlaps.numpy()
    path = os.path.join(GOLD, 'max_iou_assigner.npz')
    np.savez_compressed(path, **out)
    print(f'[golden] {path}: {os.path.getsize(path) / 1024:.0f} KiB')


def golden_rpn():
    """RPN proposal path: the real RPNHead.get_bboxes (anchor_head.py:551-590 -> rpn_head.py:78-186) with the
    real AnchorGenerator / DeltaXYWHBBoxCoder on seeded inputs vs oracle/anchors.py::rpn_proposals."""
    from mmdet.models.dense_heads.rpn_head import RPNHead

    class AttrDict(dict):                       # stands in for mmcv.ConfigDict (attribute access + .get + deepcopy)
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

    from mmdet.core.anchor import AnchorGenerator
    from oracle import anchors as oa
    c = oa.RPN_CFG
    out = {}
    ag = AnchorGenerator(scales=c['scales'], ratios=c['ratios'], strides=c['strides'])
    for l, s in enumerate(c['strides']):
        eq(oa.base_anchors(s, c['scales'], c['ratios']), ag.base_anchors[l], f'base anchors level {l}')
        eq(oa.grid_anchors(oa.base_anchors(s, c['scales'], c['ratios']), (7, 5), (s, s)),
           ag.single_level_grid_anchors(ag.base_anchors[l], (7, 5), (s, s), device='cpu'), 'grid anchors')
    vf = ag.valid_flags([(16, 20), (8, 10), (4, 5), (2, 3), (1, 2)], (60, 75, 3), device='cpu')
    out['valid_flags'] = torch.cat(vf).numpy()
    for name, seed, size, nms_pre, max_per_img in [('a', 3, (512, 640), 1000, 1000), ('b', 4, (256, 320), 300, 100), ('c', 5, (64, 96), 1000, 50)]:
        cfg = dict(c, nms_pre=nms_pre, max_per_img=max_per_img)
        test_cfg = AttrDict(dict(nms_pre=nms_pre, max_per_img=max_per_img, nms=dict(type='nms', iou_threshold=cfg['iou_threshold']),
                                  min_bbox_size=cfg['min_bbox_size']))
        head = RPNHead(in_channels=8, feat_channels=8,
                       anchor_generator=dict(type='AnchorGenerator', scales=c['scales'], ratios=c['ratios'], strides=c['strides']),
                       bbox_coder=dict(type='DeltaXYWHBBoxCoder', target_means=list(c['means']), target_stds=list(c['stds'])),
                       loss_cls=dict(type='CrossEntropyLoss', use_sigmoid=True, loss_weight=1.0), loss_bbox=dict(type='L1Loss', loss_weight=1.0),
                       test_cfg=test_cfg)
        cls, box, shapes = oa.synth_rpn_inputs(seed, size=size)
        metas = [dict(img_shape=sh, scale_factor=np.ones(4, np.float32), pad_shape=(size[0], size[1], 3)) for sh in shapes]
        with torch.no_grad():
            ref = head.get_bboxes(cls, box, metas, cfg=test_cfg, with_nms=True)
        mine, al = oa.rpn_proposals(cls, box, shapes, cfg, return_all=True)
        for b in range(len(ref)):
            eq(mine[b], ref[b], f'rpn dets {name}/{b}')
            out[f'{name}_dets{b}'] = ref[b].numpy()
            out[f'{name}_keep_pos{b}'] = al['per_image'][b]['keep_pos'].numpy().astype(np.int32)
            out[f'{name}_levels{b}'] = al['per_image'][b]['levels'].numpy().astype(np.int8)
        out[f'{name}_cand_idx'] = al['cand_idx'].numpy().astype(np.int32)
        out[f'{name}_cand_boxes_sub'] = al['cand_boxes'].reshape(-1)[::5].numpy()
    path = os.path.join(GOLD, 'rpn_proposals.npz')
    np.savez_compressed(path, **out)
    print(f'[golden] {path}: {os.path.getsize(path) / 1024:.0f} KiB')


def golden_api_signatures():
    """parameter names (in order) of the reference's plugin classes / functions on the path -> tests/golden/api_signatures.json; the
    mirrors must accept every one of them in the same relative order (tests/test_host_logic.py::test_api_signatures_match_reference)."""
    import inspect
    import json
    from mmdet.models.point.dense_heads.cpr_head import CPRHead
    from mmdet.models.point.dense_heads.p2p_head import P2PHead
    from mmdet.core.bbox.assigners import MaxIoUAssigner, PointAssigner
    from mmdet.core.bbox.assigners.hungarian_assigner import HungarianAssignerV2
    from mmdet.core.bbox.samplers import PseudoSampler
    from mmdet.core.anchor import AnchorGenerator
    from mmdet.core.post_processing.bbox_nms import multiclass_nms
    from mmdet.models.dense_heads.rpn_head import RPNHead
    from mmdet.models.detectors.base import BaseDetector

    def names(f):
        return [p.name for p in inspect.signature(f).parameters.values()
                if p.name != 'self' and p.kind not in (p.VAR_KEYWORD, p.VAR_POSITIONAL)]
    heads = ['__init__', 'forward', 'forward_train', 'simple_test', 'loss', 'get_bboxes']
    out = {}
    for cls, meths in [(CPRHead, heads), (P2PHead, heads), (MaxIoUAssigner, ['__init__', 'assign']), (PointAssigner, ['__init__', 'assign']),
                       (HungarianAssignerV2, ['__init__', 'assign']), (PseudoSampler, ['sample']),
                       (AnchorGenerator, ['__init__', 'grid_anchors', 'valid_flags', 'single_level_grid_anchors', 'gen_single_level_base_anchors']),
                       (RPNHead, ['get_bboxes'])]:
        for m in meths:
            out[f'{cls.__name__}.{m}'] = names(getattr(cls, m))
    out['multiclass_nms'] = names(multiclass_nms)
    out['BaseDetector._parse_losses'] = names(BaseDetector._parse_losses)
    path = os.path.join(GOLD, 'api_signatures.json')
    json.dump(out, open(path, 'w'), indent=1, sort_keys=True)
    print(f'[golden] {path}: {len(out)} signatures')


def golden_state_dict_keys(HEADS):
    """checkpoint compatibility: {parameter / buffer name: shape} of the REAL reference head built from every shipped CPR / P2P config
    (tests/golden/reference_head_cfgs.json) -> tests/golden/state_dict_shapes.json; the mirrors must expose exactly the same set."""
    import json
    class AttrDict(dict):                       # stands in for mmcv.ConfigDict (attribute access, nested)
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

    def wrap(x):
        if isinstance(x, dict):
            return AttrDict({k: wrap(v) for k, v in x.items()})
        if isinstance(x, (list, tuple)):
            return type(x)(wrap(v) for v in x)
        return x

    cfgs = json.load(open(os.path.join(GOLD, 'reference_head_cfgs.json')))
    out = {}
    for name, c in cfgs.items():
        if 'error' in c or c['bbox_head']['type'] == 'CascadeCPRHead':
            continue
        hc = dict(c['bbox_head'])
        head = HEADS.build(hc, default_args=dict(train_cfg=wrap(c.get('train_cfg')), test_cfg=wrap(c.get('test_cfg'))))
        out[name] = {k: list(v.shape) for k, v in head.state_dict().items()}
    path = os.path.join(GOLD, 'state_dict_shapes.json')
    json.dump(out, open(path, 'w'), indent=0, sort_keys=True)
    print(f'[golden] {path}: {len(out)} heads, {sum(len(v) for v in out.values())} entries')


def main():
    torch.set_num_threads(os.cpu_count())
    os.makedirs(GOLD, exist_ok=True)
    HEADS = load_reference()
    golden_point_assigner()
    golden_multiclass_nms_options()
    golden_cpr(HEADS, 'lite', 1234)
    golden_cpr(HEADS, 'mid', 77)
    golden_cpr(HEADS, 'lite', 99, with_towers=True)
    golden_cpr(HEADS, 'lite', 1234, grid_radius=3)
    golden_cpr(HEADS, 'mid', 77, grid_radius=2)
    for v in CPR_VARIANTS:
        golden_cpr_variant(HEADS, 'lite', 4242, v)
    golden_result_json(HEADS)
    golden_max_iou()
    golden_rpn()
    golden_api_signatures()
    golden_state_dict_keys(HEADS)
    golden_p2p(HEADS, 'lite', 4321, 0.01)
    golden_p2p(HEADS, 'mid', 555, 0.5)
    golden_p2p(HEADS, 'mid', 555, 0.01)
    golden_p2p_aug(HEADS, 'lite', 2468, 0.5)
    golden_p2p_aug(HEADS, 'mid', 1357, 0.3)


if __name__ == '__main__':
    main()
