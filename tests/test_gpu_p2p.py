"""GPU parity of the P2P path: decode/top-k, multiclass NMS, cost matrix, assigner, losses, head API."""
import os

import numpy as np
import pytest
import torch

from oracle import p2p as op2p, synth
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from pointtinybenchmark_b200 import ops
    return ops


def head_cfg(d, nms_iou=0.01):
    return dict(
        type='P2PHead', norm_cfg=dict(type='GN', num_groups=32, requires_grad=True), num_classes=d['num_classes'],
        in_channels=d['C'], feat_channels=d['C'], stacked_convs=4, strides=[d['stride']], point_anchor=[(0., 0.)],
        loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0),
        loss_reg=dict(type='SmoothL1Loss', beta=1.0 / 9.0, loss_weight=0.5), pts_gamma=1, reg_norm=1,
        train_cfg=dict(neg_weight=1.0, assigner=dict(type='HungarianAssignerV2', cls_costs=dict(type='FocalLossCost', weight=2.0),
                                                     reg_costs=dict(type='DisCostV2', weight=0.1, norm_with_img_wh=False),
                                                     topk_k=5), sampler=dict(type='PseudoSampler')),
        test_cfg=dict(nms_pre=1000, min_bbox_size=0, score_thr=0.05, pseudo_wh=(32, 32),
                      nms=dict(type='nms', iou_threshold=nms_iou), max_per_img=100))


@pytest.mark.parametrize('name,seed,iou', [('lite', 4321, 0.01), ('mid', 555, 0.5), ('mid', 555, 0.01)])
def test_get_bboxes_topk_and_nms_bit_exact(ops, golden_dir, name, seed, iou):
    from pointtinybenchmark_b200.p2p_head import P2PHead  # noqa
    from pointtinybenchmark_b200.registry import build_head
    dev = torch.device('cuda:0')
    inp = synth.p2p_inputs(name, seed)
    d = inp['cfgd']
    cfg = op2p.default_cfg(num_classes=d['num_classes'], stride=d['stride'], nms_iou=iou)
    gold = np.load(os.path.join(golden_dir, f'p2p_{name}_iou{iou}.npz'))
    head = build_head(head_cfg(d, iou)).cuda().eval()
    cls_out, pts_out = inp['cls_out'].to(dev), inp['pts_out'].to(dev)
    res, aux = head.get_bboxes([cls_out], [pts_out], inp['img_metas'], return_all=True)
    _, pred, _, cls = op2p.pred_points(inp['cls_out'], inp['pts_out'], inp['img_metas'], cfg)
    o_topk, o_keep, o_cand = [], [], []
    for b, m in enumerate(inp['img_metas']):
        ps, labels, al = op2p.get_bboxes_single(pred[b][..., :2], cls[b], m['img_shape'], m['scale_factor'], cfg, return_all=True)
        # ---- integer outputs: bit exact
        assert torch.equal(aux['topk_idx'][b].cpu().long(), al['topk_inds']), f'top-k indices image {b}'
        assert int(aux['cand_count'][b]) == len(al['cand_inds'])
        n = int(aux['count'][b])
        assert n == len(al['keep'])
        assert torch.equal(aux['keep'][b, :n].cpu().long(), al['keep']), f'NMS keep indices image {b}'
        assert torch.equal(res[b][1].cpu(), labels), 'labels'
        # ---- float outputs
        wh = torch.tensor(cfg['pseudo_wh'])
        ref_boxes = torch.cat([ps[:, :2] - wh / 2, ps[:, :2] + wh / 2, ps[:, 2:]], -1)
        assert_close(res[b][0], ref_boxes, 1e-4, f'pseudo boxes image {b}')
        assert_close(aux['scores'][b], al['scores'], 1e-4, 'top-k scores')
        o_topk.append(al['topk_inds']); o_keep.append(al['keep']); o_cand.append(len(al['cand_inds']))
    assert np.array_equal(torch.cat(o_topk).numpy().astype(np.int32), gold['topk'])
    assert np.array_equal(torch.cat(o_keep).numpy(), gold['keep'])
    got_keep = torch.cat([aux['keep'][b, :int(aux['count'][b])] for b in range(len(res))]).cpu().numpy()
    assert np.array_equal(got_keep.astype(np.int64), gold['keep']), 'keep vs golden'
    got_topk = aux['topk_idx'].cpu().numpy().reshape(-1)
    assert np.array_equal(got_topk, gold['topk']), 'topk vs golden'
    assert_close(torch.cat([r[0] for r in res]), torch.from_numpy(gold['det']), 1e-4, 'det vs golden')


def test_nms_edge_cases(ops):
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(3)
    # (a) nothing above threshold; (b) nms_pre >= Q identity; (c) heavy overlap, single class
    B, P, C = 2, 300, 5
    pts = (torch.rand(B, P, 2, generator=g) * 60).contiguous()
    scores = torch.rand(B, P, C, generator=g) * 0.04
    cnt, det, lab, keep, cc = ops.multiclass_nms(pts.to(dev), scores.to(dev), (32, 32), 0.05, 0.01, 100)
    assert cnt.tolist() == [0, 0] and cc.tolist() == [0, 0]
    scores = torch.rand(B, P, C, generator=g)
    scores += torch.arange(scores.numel()).reshape(scores.shape).float() * 1e-7
    for iou, mx in [(0.01, 100), (0.5, 7), (0.9, 1000)]:
        cnt, det, lab, keep, cc = ops.multiclass_nms(pts.to(dev), scores.to(dev), (32, 32), 0.3, iou, mx)
        for b in range(B):
            boxes = torch.cat([pts[b] - 16, pts[b] + 16], -1)
            sb = torch.cat([scores[b], scores[b].new_zeros(P, 1)], 1)
            d, l, k, inds = op2p.multiclass_nms(boxes, sb, 0.3, iou, mx)
            n = int(cnt[b])
            assert n == len(k) and int(cc[b]) == len(inds)
            assert torch.equal(keep[b, :n].cpu().long(), k), (iou, mx, b)
            assert torch.equal(lab[b, :n].cpu().long(), l)
            assert torch.equal(det[b, :n].cpu(), d), 'dets are copies of the inputs: exact'
    # torchvision cross-check of the oracle itself (same IoU>thr / offset-0 semantics)
    import torchvision
    b = torch.cat([pts[0] - 16, pts[0] + 16], -1)
    s = scores[0, :, 0].contiguous()
    assert torch.equal(op2p.nms(b, s, 0.3), torchvision.ops.nms(b, s, 0.3))


def test_decode_identity_when_nms_pre_exceeds(ops):
    dev = torch.device('cuda:0')
    inp = synth.p2p_inputs('lite', 11)
    d = inp['cfgd']
    cfg = op2p.default_cfg(num_classes=d['num_classes'], stride=d['stride'])
    H, W = inp['cls_out'].shape[2:]
    cm = ops.to_nhwc(inp['cls_out'].to(dev)).contiguous()
    rm = ops.to_nhwc(inp['pts_out'].to(dev)).contiguous()
    img_hw = torch.tensor([m['img_shape'][:2] for m in inp['img_metas']], dtype=torch.int32, device=dev)
    anchor = torch.tensor([[0., 0.]], device=dev)
    idx, pts, sc = ops.p2p_decode_topk(cm, rm, d['num_classes'], 1, anchor, d['stride'], 1.0, img_hw, -1)
    assert idx.shape[1] == H * W and torch.equal(idx[0].cpu(), torch.arange(H * W, dtype=torch.int32))
    _, pred, _, cls = op2p.pred_points(inp['cls_out'], inp['pts_out'], inp['img_metas'], cfg)
    x = pred[0][:, 0].clamp(0, inp['img_metas'][0]['img_shape'][1])
    y = pred[0][:, 1].clamp(0, inp['img_metas'][0]['img_shape'][0])
    assert torch.equal(pts[0].cpu(), torch.stack([x, y], -1)), 'decoded points must be bit-identical (same fp32 op order)'
    assert_close(sc[0], cls[0].sigmoid(), 1e-5, 'scores')
    # 4 anchors per cell
    g = torch.Generator().manual_seed(0)
    k, C = 4, 8
    cls4 = (torch.randn(1, k * C, 9, 11, generator=g) * 2).to(dev)
    reg4 = torch.randn(1, 2 * k, 9, 11, generator=g).to(dev)
    pa = [(-0.25, -0.25), (0.25, -0.25), (0.25, 0.25), (-0.25, 0.25)]
    cfg4 = op2p.default_cfg(num_classes=C, stride=8, point_anchor=pa, pts_gamma=100. / 8, nms_pre=50)
    metas = [dict(pad_shape=(72, 88, 3), img_shape=(70, 85, 3), scale_factor=[1., 1., 1., 1.])]
    _, pred, _, cls = op2p.pred_points(cls4.cpu(), reg4.cpu(), metas, cfg4)
    _, _, al = op2p.get_bboxes_single(pred[0][..., :2], cls[0], metas[0]['img_shape'], metas[0]['scale_factor'], cfg4, return_all=True)
    idx, pts, sc = ops.p2p_decode_topk(ops.to_nhwc(cls4).contiguous(), ops.to_nhwc(reg4).contiguous(), C, k,
                                       torch.tensor(pa, device=dev), 8, 100. / 8,
                                       torch.tensor([[70, 85]], dtype=torch.int32, device=dev), 50)
    assert torch.equal(idx[0].cpu().long(), al['topk_inds'])
    assert_close(sc[0], al['scores'], 1e-5, 'scores k=4')


@pytest.mark.parametrize('name,seed', [('lite', 4321), ('mid', 555)])
def test_loss_targets_and_grads(ops, golden_dir, name, seed):
    from pointtinybenchmark_b200.registry import build_head
    from pointtinybenchmark_b200 import p2p_head  # noqa
    dev = torch.device('cuda:0')
    inp = synth.p2p_inputs(name, seed)
    d = inp['cfgd']
    cfg = op2p.default_cfg(num_classes=d['num_classes'], stride=d['stride'])
    gold = np.load(os.path.join(golden_dir, f'p2p_{name}_iou0.01.npz'))
    head = build_head(head_cfg(d)).cuda()
    co = inp['cls_out'].to(dev).requires_grad_(True)
    po = inp['pts_out'].to(dev).requires_grad_(True)
    gtb = [b.to(dev) for b in inp['gt_bboxes']]
    gtl = [l.to(dev) for l in inp['gt_labels']]
    losses = head.loss([co], [po], gtb, gtl, inp['img_metas'])
    (sum(losses['loss_cls']) + sum(losses['loss_pts'])).backward()
    co_o = inp['cls_out'].clone().requires_grad_(True)
    po_o = inp['pts_out'].clone().requires_grad_(True)
    ol, oall = op2p.p2p_loss(co_o, po_o, inp['gt_bboxes'], inp['gt_labels'], inp['img_metas'], cfg, return_all=True)
    (sum(ol['loss_cls']) + sum(ol['loss_pts'])).backward()
    # assignment (int) must agree with the oracle's scipy run and the golden
    for b in range(d['B']):
        assert torch.equal(head._last_targets['labels'][b].cpu(), oall['targets'][b][0]), 'assigned labels'
        assert torch.equal(head._last_targets['pts_weights'][b].cpu(), oall['targets'][b][3])
    for k in ('loss_cls', 'loss_pts'):
        assert_close(torch.stack(losses[k]), torch.stack(ol[k]).detach(), 1e-4, k)
        assert_close(torch.stack(losses[k]), torch.from_numpy(gold[k]), 1e-4, k + ' vs golden')
    assert_close(co.grad, co_o.grad, 2e-4, 'd/d cls_out')
    assert_close(po.grad, po_o.grad, 2e-4, 'd/d pts_out')
    # cost matrix kernel vs oracle (first image)
    _, pred, valid, cls = op2p.pred_points(inp['cls_out'], inp['pts_out'], inp['img_metas'], cfg)
    gp = (inp['gt_bboxes'][0][:, :2] + inp['gt_bboxes'][0][:, 2:]) / 2
    ref = op2p.cost_matrix(pred[0][valid[0]][..., :2], cls[0][valid[0]], gp, inp['gt_labels'][0], inp['img_metas'][0]['img_shape'], cfg)
    ridx = torch.nonzero(valid[0]).squeeze(1).int().to(dev)
    got = ops.p2p_cost_matrix(cls[0].to(dev).contiguous(), pred[0][:, :2].contiguous().to(dev), ridx, gp.to(dev).contiguous(),
                              inp['gt_labels'][0].int().to(dev), 2.0, 0.25, 2, 1e-12, 0.1)
    assert_close(got, ref, 1e-5, 'cost matrix')
    assert np.abs(got.cpu().flatten()[::37].numpy() - gold['cost_sub']).max() <= 1e-5 * np.abs(gold['cost_sub']).max()


def test_point_assigner_reference_kats(ops, golden_dir):
    """the reference's own golden vectors: TOV_mmdetection/tests/test_utils/test_assigner.py:155-194"""
    dev = torch.device('cuda:0')
    pts = torch.FloatTensor([[0, 0, 1], [10, 10, 1], [5, 5, 1], [32, 32, 1]]).to(dev)
    gts = torch.FloatTensor([[0, 0, 10, 9], [0, 10, 10, 19]]).to(dev)
    assert ops.point_assigner(pts, gts).tolist() == [1, 2, 1, 0]
    assert ops.point_assigner(pts, torch.zeros(0, 4, device=dev)).tolist() == [0, 0, 0, 0]
    assert len(ops.point_assigner(torch.zeros(0, 3, device=dev), torch.zeros(0, 4, device=dev))) == 0
    gold = np.load(os.path.join(golden_dir, 'point_assigner.npz'))
    got = ops.point_assigner(torch.from_numpy(gold['points']).to(dev), torch.from_numpy(gold['gts']).to(dev), 4, 3)
    assert np.array_equal(got.cpu().numpy(), gold['gt_inds'])


def test_nms_on_explicit_boxes_and_tta_merge(ops):
    """second NMS of the test-time-aug path: arbitrary boxes (different sizes), class-offset semantics, vs the oracle."""
    from pointtinybenchmark_b200.p2p_head import P2PHead
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(21)
    P, C = 500, 7
    c = torch.rand(P, 2, generator=g) * torch.tensor([400., 260.])
    wh = torch.rand(P, 2, generator=g) * 40 + 6
    boxes = torch.cat([c - wh / 2, c + wh / 2], 1)
    scores = torch.rand(P, C, generator=g) * (torch.rand(P, C, generator=g) > 0.7).float()
    scores += (scores > 0).float() * torch.arange(P * C).reshape(P, C).float() * 1e-7
    for iou, mx in [(0.5, 100), (0.2, 30)]:
        cnt, det, lab, keep, cc = ops.multiclass_nms_boxes(boxes[None].to(dev), scores[None].to(dev), 0.05, iou, mx)
        d, l, k, inds = op2p.multiclass_nms(boxes, torch.cat([scores, scores.new_zeros(P, 1)], 1), 0.05, iou, mx)
        n = int(cnt[0])
        assert n == len(k) and int(cc[0]) == len(inds)
        assert torch.equal(keep[0, :n].cpu().long(), k) and torch.equal(lab[0, :n].cpu().long(), l)
        assert torch.equal(det[0, :n].cpu(), d)
    # bbox_mapping_back incl. flip + tile offset (core/bbox/transforms.py:62-85)
    b = torch.tensor([[10., 20., 30., 50.]])
    out = P2PHead.bbox_mapping_back(b, (100, 200, 3), [2., 2., 2., 2.], True, 'horizontal', (5, 7))
    assert torch.equal(out, torch.tensor([[(200 - 30) / 2 + 5, 10. + 7, (200 - 10) / 2 + 5, 25. + 7]]))


@pytest.mark.parametrize('method,iou,sigma,min_score', [('linear', 0.3, 0.5, 1e-3), ('gaussian', 0.5, 0.5, 0.05), ('naive', 0.3, 0.5, 1e-3),
                                                        ('gaussian', 0.3, 0.25, 0.02)])
def test_soft_nms_matches_the_restated_mmcv_algorithm(ops, method, iou, sigma, min_score):
    """multiclass soft-NMS (nms=dict(type='soft_nms', ...)): keep indices, labels and selection order are integer targets, the decayed
    scores float.  Oracle = the sequential mmcv CPU loop restated in oracle/p2p.py::soft_nms, run on the offset boxes of ALL classes
    at once (what batched_nms does); the kernel decomposes per class."""
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(17)
    B, P, C = 3, 400, 7
    pts = (torch.rand(B, P, 2, generator=g) * torch.tensor([300.0, 200.0])).contiguous()
    scores = (torch.rand(B, P, C, generator=g) ** 6).contiguous()                    # ~12 % of the (point, class) pairs above 0.05
    scores += torch.arange(B * P * C).reshape(B, P, C) * 1e-9                          # distinct scores
    wh, thr, max_n = (32.0, 32.0), 0.05, 100
    cnt, det, lab, keep, cc = ops.multiclass_soft_nms(pts.to(dev), scores.to(dev), wh, thr, iou, max_n, sigma, min_score, method)
    nms_cfg = dict(type='soft_nms', iou_threshold=iou, sigma=sigma, min_score=min_score, method=method)
    half = torch.tensor(wh) / 2
    for b in range(B):
        boxes = torch.cat([pts[b] - half, pts[b] + half], -1)
        ms = torch.cat([scores[b], torch.zeros(P, 1)], -1)
        o_det, o_lab, o_keep, o_inds = op2p.multiclass_nms(boxes, ms, thr, iou, max_n, nms_cfg=nms_cfg)
        n = int(cnt[b])
        assert int(cc[b]) == len(o_inds) and n == len(o_keep), (method, b, n, len(o_keep))
        assert torch.equal(keep[b, :n].cpu().long(), o_keep), f'{method}: keep indices / selection order, image {b}'
        assert torch.equal(lab[b, :n].cpu().long(), o_lab)
        assert torch.equal(det[b, :n, :4].cpu(), o_det[:, :4])
        assert_close(det[b, :n, 4], o_det[:, 4], 1e-5, 'decayed scores')
        assert bool((det[b, :n - 1, 4] >= det[b, 1:n, 4]).all())
    # through the head: P2PHead with nms type soft_nms == the op
    from pointtinybenchmark_b200 import p2p_head  # noqa: F401  (registers the head)
    from pointtinybenchmark_b200.registry import build_head
    inp = synth.p2p_inputs('lite', 4321)
    hc = head_cfg(inp['cfgd'], iou)
    hc['test_cfg']['nms'] = nms_cfg
    head = build_head(hc).cuda().eval()
    res = head.get_bboxes([inp['cls_out'].to(dev)], [inp['pts_out'].to(dev)], inp['img_metas'])
    cfg = op2p.default_cfg(num_classes=inp['cfgd']['num_classes'], stride=inp['cfgd']['stride'], nms_iou=iou)
    _, pred, _, cls = op2p.pred_points(inp['cls_out'], inp['pts_out'], inp['img_metas'], cfg)
    for b, m in enumerate(inp['img_metas']):
        sc = cls[b].sigmoid()
        _, topk = sc.max(dim=1)[0].topk(min(cfg['nms_pre'], sc.shape[0]))
        p = pred[b][topk][:, :2]
        p = torch.stack([p[:, 0].clamp(0, m['img_shape'][1]), p[:, 1].clamp(0, m['img_shape'][0])], -1)
        whp = torch.tensor(cfg['pseudo_wh']) / 2
        o_det, o_lab, _, _ = op2p.multiclass_nms(torch.cat([p - whp, p + whp], -1), torch.cat([sc[topk], torch.zeros(len(topk), 1)], -1),
                                                 cfg['score_thr'], iou, cfg['max_per_img'], nms_cfg=nms_cfg)
        assert torch.equal(res[b][1].cpu(), o_lab)
        assert_close(res[b][0], o_det, 1e-4, 'soft-NMS detections through P2PHead.get_bboxes')


@pytest.mark.parametrize('kind', ['nms', 'linear', 'gaussian'])
def test_class_offset_corner_case_takes_the_exact_global_path(ops, kind):
    """batched_nms separates classes by adding label*(max_coord+1): with NEGATIVE coordinates (pseudo boxes of points near the top-left
    corner) on a square image, a class c box in the negative corner intersects a class c-1 box near max_coord.  Those images must
    reproduce the reference's all-classes-at-once loop (hard NMS: nms_global_kernel, soft NMS: soft_nms_global_kernel)."""
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(5)
    P, C = 200, 4
    pts = torch.rand(1, P, 2, generator=g) * 240 + 8
    pts[0, :6] = torch.tensor([[0., 0.], [1., 2.], [2., 1.], [255., 255.], [254., 256.], [256., 254.]])   # both extreme corners
    scores = torch.rand(1, P, C, generator=g) ** 4 + torch.arange(P * C).reshape(1, P, C) * 1e-9
    scores[0, :6] = 0.5 + torch.rand(6, C, generator=g) * 0.4                                              # corner points are candidates
    wh, thr, iou, max_n = (32.0, 32.0), 0.05, 0.01, 100
    half = torch.tensor(wh) / 2
    boxes = torch.cat([pts[0] - half, pts[0] + half], -1)
    ms = torch.cat([scores[0], torch.zeros(P, 1)], -1)
    # the offset really fails to separate classes here: some cross-class pair of offset boxes intersects
    lab = torch.arange(C).repeat(P)
    bb = boxes.repeat_interleave(C, 0) + (lab.float() * (boxes.max() + 1))[:, None]
    inter = (torch.min(bb[:, None, 2:], bb[None, :, 2:]) - torch.max(bb[:, None, :2], bb[None, :, :2])).clamp(min=0).prod(-1)
    assert bool((inter[lab[:, None] != lab[None, :]] > 0).any()), 'fixture must contain a cross-class intersection'
    if kind == 'nms':
        cnt, det, labels, keep, cc = ops.multiclass_nms(pts.to(dev), scores.to(dev), wh, thr, iou, max_n)
        o_det, o_lab, o_keep, _ = op2p.multiclass_nms(boxes, ms, thr, iou, max_n)
    else:
        cfg = dict(type='soft_nms', iou_threshold=0.3, sigma=0.5, min_score=1e-3, method=kind)
        cnt, det, labels, keep, cc = ops.multiclass_soft_nms(pts.to(dev), scores.to(dev), wh, thr, 0.3, max_n, 0.5, 1e-3, kind)
        o_det, o_lab, o_keep, _ = op2p.multiclass_nms(boxes, ms, thr, 0.3, max_n, nms_cfg=cfg)
    n = int(cnt[0])
    assert n == len(o_keep)
    assert torch.equal(keep[0, :n].cpu().long(), o_keep) and torch.equal(labels[0, :n].cpu().long(), o_lab)
    assert_close(det[0, :n], o_det, 1e-5, 'detections')


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['factors', 'class_specific', 'agnostic', 'unlimited'])
def test_multiclass_nms_mirror_options_vs_reference_golden(name, golden_dir):
    """`post_processing.multiclass_nms` with the options the point heads never pass (score_factors, class-specific boxes, class_agnostic,
    max_num=-1) against the outputs of the REAL reference function (tests/golden/multiclass_nms_options.npz): keep indices and labels
    bit-exact, dets exact (boxes are copies, scores one fp32 product)."""
    import os
    from pointtinybenchmark_b200.post_processing import multiclass_nms
    from tests.test_oracle_golden import NMS_OPTION_CASES
    c = NMS_OPTION_CASES[name]
    z = np.load(os.path.join(golden_dir, 'multiclass_nms_options.npz'))
    dev = torch.device('cuda:0')
    sf = torch.from_numpy(z['factors']).to(dev) if c['sf'] else None
    d, l, k = multiclass_nms(torch.from_numpy(z[c['boxes']]).to(dev), torch.from_numpy(z['scores']).to(dev), 0.05, dict(c['cfg']), c['max_num'],
                             score_factors=sf, return_inds=True)
    assert np.array_equal(k.cpu().numpy(), z[f'{name}_keep'])
    assert np.array_equal(l.cpu().numpy(), z[f'{name}_labels'])
    assert np.array_equal(d.cpu().numpy(), z[f'{name}_dets'])
