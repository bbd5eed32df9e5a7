"""CPU: the oracle restatement reproduces the golden vectors that oracle/make_golden.py recorded from the REAL reference."""
import os

import numpy as np
import pytest
import torch

from oracle import cpr as ocpr, p2p as op2p, synth
from tests.helpers import oracle_cfg


def _sub(t, step):
    return t.detach().flatten()[::step].numpy()


@pytest.mark.parametrize('name,seed', [('lite', 1234), ('mid', 77)])
def test_cpr_oracle_matches_reference_golden(golden_dir, name, seed):
    inp = synth.cpr_inputs(name, seed)
    cfg = oracle_cfg(inp['cfgd'])
    gold = np.load(os.path.join(golden_dir, f'cpr_{name}.npz'))
    assert int(gold['seed']) == seed
    gt_r = [p.reshape(len(l), -1, 2) for p, l in zip(ocpr.pseudo_bbox_to_center(inp['gt_bboxes']), inp['gt_labels'])]
    ex = ocpr.extract(inp['cls_feat'], gt_r, inp['gt_labels'], inp['img_metas'], cfg)
    assert np.array_equal(ex['pos_valid'].numpy(), gold['pos_valid'])
    assert np.array_equal(ex['pos_pts'].numpy(), gold['pos_pts'])
    gneg = np.unpackbits(gold['neg_valid'])[:int(np.prod(gold['neg_valid_shape']))].reshape(gold['neg_valid_shape'])
    assert np.array_equal(ex['neg_valid'].numpy(), gneg.astype(bool))
    assert np.array_equal(_sub(ex['pos_feats'], 1009), gold['pos_feats_sub'])
    losses, al = ocpr.cpr_loss(inp['cls_feat'], inp['weights'], inp['gt_bboxes'], inp['gt_labels'], inp['img_metas'], cfg,
                               return_all=True)
    for k in ('gt_loss', 'pos_loss', 'neg_loss', 'bag_acc'):
        np.testing.assert_allclose(losses[k].reshape(-1).numpy(), gold['loss_' + k], rtol=1e-6)
    np.testing.assert_allclose(al['bag_prob'].numpy(), gold['mil_bag_prob'], rtol=1e-6, atol=1e-9)
    res, ra = ocpr.cpr_get_bboxes(inp['cls_feat'], inp['weights'], inp['gt_bboxes'], inp['gt_labels'], inp['gt_anns_id'],
                                  inp['img_metas'], cfg, return_all=True)
    assert np.array_equal(torch.cat([r[0] for r in res]).numpy(), gold['det'])
    for key in ('not_refine', 'chosen', 'merge_valid', 'mask_nearest', 'mask_classify'):
        assert np.array_equal(torch.cat([r[key] for r in ra['refine']]).numpy(), gold[key]), key
    assert 0.05 < float(gold['frac_not_refine']) < 0.6, 'fixture must exercise both the refine and the fallback path'


@pytest.mark.parametrize('name,seed,iou', [('lite', 4321, 0.01), ('mid', 555, 0.5), ('mid', 555, 0.01)])
def test_p2p_oracle_matches_reference_golden(golden_dir, name, seed, iou):
    inp = synth.p2p_inputs(name, seed)
    d = inp['cfgd']
    cfg = op2p.default_cfg(num_classes=d['num_classes'], stride=d['stride'], nms_iou=iou)
    gold = np.load(os.path.join(golden_dir, f'p2p_{name}_iou{iou}.npz'))
    _, pred, valid, cls = op2p.pred_points(inp['cls_out'], inp['pts_out'], inp['img_metas'], cfg)
    dets, keeps, topks = [], [], []
    for b, m in enumerate(inp['img_metas']):
        ps, labels, al = op2p.get_bboxes_single(pred[b][..., :2], cls[b], m['img_shape'], m['scale_factor'], cfg, return_all=True)
        wh = torch.tensor(cfg['pseudo_wh'])
        dets.append(torch.cat([ps[:, :2] - wh / 2, ps[:, :2] + wh / 2, ps[:, 2:]], -1))
        keeps.append(al['keep']); topks.append(al['topk_inds'])
    assert np.array_equal(torch.cat(keeps).numpy(), gold['keep'])
    assert np.array_equal(torch.cat(topks).numpy().astype(np.int32), gold['topk'])
    assert np.array_equal(torch.cat(dets).numpy(), gold['det'])
    losses, al = op2p.p2p_loss(inp['cls_out'], inp['pts_out'], inp['gt_bboxes'], inp['gt_labels'], inp['img_metas'], cfg,
                               return_all=True)
    np.testing.assert_allclose(torch.stack(losses['loss_cls']).numpy(), gold['loss_cls'], rtol=1e-6)
    np.testing.assert_allclose(torch.stack(losses['loss_pts']).numpy(), gold['loss_pts'], rtol=1e-6)
    assert np.array_equal(torch.stack([t[4] for t in al['targets']]).numpy().astype(np.int32), gold['gt_inds'])


@pytest.mark.parametrize('name,seed,iou', [('lite', 2468, 0.5), ('mid', 1357, 0.3)])
def test_p2p_aug_test_oracle_matches_reference_golden(golden_dir, name, seed, iou):
    """P2PHead.aug_test_bboxes (p2p_head.py:487-572): flip / scale / tile_offset mapping + the second multiclass NMS."""
    inp = synth.p2p_aug_inputs(name, seed)
    d = inp['cfgd']
    cfg = op2p.default_cfg(num_classes=d['num_classes'], stride=d['stride'], nms_iou=iou)
    gold = np.load(os.path.join(golden_dir, f'p2p_aug_{name}.npz'))
    for rescale in (False, True):
        res, aux = op2p.aug_test_bboxes(inp['outs'], inp['metas'], cfg, rescale=rescale)
        assert np.array_equal(res[0][0].numpy(), gold[f'det_rescale{int(rescale)}'])
        assert np.array_equal(res[0][1].numpy(), gold[f'labels_rescale{int(rescale)}'])
    assert np.array_equal(aux['keep'].numpy(), gold['keep'])
    assert len(aux['merged_boxes']) == int(gold['n_merged']) > len(gold['keep']), 'the second NMS must have something to suppress'


def test_point_assigner_reference_kats(golden_dir):
    """golden vectors of the reference's own test-suite: TOV_mmdetection/tests/test_utils/test_assigner.py:155-194"""
    pts = torch.FloatTensor([[0, 0, 1], [10, 10, 1], [5, 5, 1], [32, 32, 1]])
    gts = torch.FloatTensor([[0, 0, 10, 9], [0, 10, 10, 19]])
    assert op2p.point_assigner(pts, gts).tolist() == [1, 2, 1, 0]
    assert op2p.point_assigner(pts, torch.zeros(0, 4)).tolist() == [0, 0, 0, 0]
    assert len(op2p.point_assigner(torch.zeros(0, 3), torch.zeros(0, 4))) == 0
    gold = np.load(os.path.join(golden_dir, 'point_assigner.npz'))
    assert np.array_equal(op2p.point_assigner(torch.from_numpy(gold['points']), torch.from_numpy(gold['gts'])).numpy(), gold['gt_inds'])


def test_nms_oracle_against_torchvision():
    """third-party mmcv NMS semantics (sort desc, IoU > thr, offset 0) pinned to torchvision's CPU kernel."""
    import torchvision
    g = torch.Generator().manual_seed(1)
    for n, thr in [(1, 0.5), (50, 0.01), (400, 0.3), (400, 0.7)]:
        c = torch.rand(n, 2, generator=g) * 100
        wh = torch.rand(n, 2, generator=g) * 30 + 2
        boxes = torch.cat([c - wh / 2, c + wh / 2], 1)
        scores = torch.rand(n, generator=g) + torch.arange(n) * 1e-6
        assert torch.equal(op2p.nms(boxes, scores, thr), torchvision.ops.nms(boxes, scores, thr))
    assert len(op2p.multiclass_nms(torch.zeros(3, 4), torch.zeros(3, 5), 0.05, 0.5, 10)[0]) == 0


def test_soft_nms_oracle_properties():
    """mmcv's soft_nms is third-party and absent (parity unpinned): the restatement is checked through properties that follow from
    the published algorithm — 'naive' keeps exactly the hard-NMS set when min_score -> 0 (>= vs > only matters at IoU == thr), scores
    come out non-increasing, the top-scoring box is untouched, gaussian never removes a box unless its decayed score < min_score."""
    g = torch.Generator().manual_seed(0)
    c = torch.rand(80, 2, generator=g) * 100
    b = torch.cat([c - 16, c + 16], 1)
    s = torch.rand(80, generator=g)
    d, k = op2p.soft_nms(b, s, 0.3, 0.5, 1e-9, 'naive')
    assert sorted(k.tolist()) == sorted(op2p.nms(b, s, 0.3).tolist())
    for m in ('naive', 'linear', 'gaussian'):
        d, k = op2p.soft_nms(b, s, 0.3, 0.5, 0.05, m)
        assert bool((d[:-1, 4] >= d[1:, 4]).all()) and int(k[0]) == int(s.argmax()) and float(d[0, 4]) == float(s.max())
        assert torch.equal(d[:, :4], b[k]) and len(set(k.tolist())) == len(k)
        assert bool((d[:, 4] >= 0.05).all()) or m == 'naive'
    d, k = op2p.soft_nms(b, s, 0.3, 0.5, 0.0, 'gaussian')
    assert len(k) == 80                                   # nothing can drop below 0
    dets, labels, keep, inds = op2p.multiclass_nms(b, torch.cat([s[:, None], (1 - s)[:, None], torch.zeros(80, 1)], 1), 0.05, 0.3, 20,
                                                   nms_cfg=dict(type='soft_nms', iou_threshold=0.3, method='linear'))
    assert len(keep) == 20 and bool((dets[:-1, 4] >= dets[1:, 4]).all()) and set(labels.tolist()) <= {0, 1}


@pytest.mark.parametrize('variant', ['softmax', 'normed_sigmoid', 'fcs_binary'])
def test_cpr_variant_oracle_matches_reference_golden(golden_dir, variant):
    """non-default CPRHead variants (prob_cls_type softmax / normed_sigmoid, num_cls_fcs + binary_ins, gt_weights): the oracle vs the
    vectors the REAL reference head produced (oracle/make_golden.py::golden_cpr_variant)."""
    from oracle.make_golden import CPR_VARIANTS, variant_weights
    inp = synth.cpr_inputs('lite', 4242)
    d = inp['cfgd']
    gold = np.load(os.path.join(golden_dir, f'cpr_lite_{variant}.npz'))
    cfg = ocpr.default_cfg(num_classes=d['num_classes'], in_channels=d['C'], feat_channels=d['C'], stride=d['stride'], pos_radius=d['radius'],
                           neg_radius=d['radius'], **CPR_VARIANTS[variant][1])
    w = variant_weights(inp, variant, 4242)
    g = torch.Generator().manual_seed(4242 + 5)
    gtw = [torch.rand(len(l), generator=g) * 0.5 + 0.5 for l in inp['gt_labels']]
    gtw[0][0] = 0.0
    ol = ocpr.cpr_loss(inp['cls_feat'], w, inp['gt_bboxes'], inp['gt_labels'], inp['img_metas'], cfg, gt_weights=gtw)
    for k in ('gt_loss', 'pos_loss', 'neg_loss', 'bag_acc'):
        np.testing.assert_allclose(ol[k].detach().reshape(-1).numpy(), gold['loss_' + k], rtol=1e-6)
    res, ra = ocpr.cpr_get_bboxes(inp['cls_feat'], w, inp['gt_bboxes'], inp['gt_labels'], inp['gt_anns_id'], inp['img_metas'], cfg, return_all=True)
    assert np.array_equal(torch.cat([r[0] for r in res]).numpy(), gold['det'])
    assert np.array_equal(torch.cat([r['not_refine'] for r in ra['refine']]).numpy(), gold['not_refine'])


NMS_OPTION_CASES = dict(factors=dict(boxes='boxes', sf=True, cfg=dict(type='nms', iou_threshold=0.5), max_num=50),
                        class_specific=dict(boxes='boxes_cs', sf=False, cfg=dict(type='nms', iou_threshold=0.4), max_num=80),
                        agnostic=dict(boxes='boxes', sf=False, cfg=dict(type='nms', iou_threshold=0.5, class_agnostic=True), max_num=60),
                        unlimited=dict(boxes='boxes', sf=True, cfg=dict(type='nms', iou_threshold=0.3), max_num=-1))


def test_multiclass_nms_options_match_reference_golden(golden_dir):
    """bbox_nms.py:7-94 with score_factors / class-specific boxes / class_agnostic / max_num=-1: the oracle restatement against the
    outputs recorded from the REAL reference function (oracle/make_golden.py::golden_multiclass_nms_options)."""
    import os
    from oracle import p2p as op2p
    z = np.load(os.path.join(golden_dir, 'multiclass_nms_options.npz'))
    scores = torch.from_numpy(z['scores'])
    for name, c in NMS_OPTION_CASES.items():
        sf = torch.from_numpy(z['factors']) if c['sf'] else None
        d, l, k, _ = op2p.multiclass_nms(torch.from_numpy(z[c['boxes']]), scores, 0.05, c['cfg']['iou_threshold'], c['max_num'], nms_cfg=c['cfg'],
                                         score_factors=sf)
        assert np.array_equal(k.numpy(), z[f'{name}_keep']), name
        assert np.array_equal(l.numpy(), z[f'{name}_labels']), name
        assert np.array_equal(d.numpy(), z[f'{name}_dets']), name
