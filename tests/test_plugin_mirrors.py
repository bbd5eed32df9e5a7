"""Host logic of the assigner / sampler / NMS plugin mirrors (pointtinybenchmark_b200/{assigners,post_processing,registry}.py): the glue
around the C ABI is exercised on CPU by standing the ORACLE in for the library calls (the kernels themselves are compared with the same
oracle in the -m gpu tests), so argument handling, result structures and the reference's own test vectors
(TOV_mmdetection/tests/test_utils/test_assigner.py:155-194) are checked without a GPU."""
import numpy as np
import pytest
import torch

from oracle import anchors as oa
from oracle import p2p as op2p
from pointtinybenchmark_b200 import assigners, ops, post_processing, registry


class FakeCuda(torch.Tensor):
    """a CPU tensor that reports is_cuda: lets the mirrors' CUDA-only guards pass in this dry run"""
    is_cuda = property(lambda s: True)


def fc(t):
    return t.as_subclass(FakeCuda)


@pytest.fixture()
def oracle_backend(monkeypatch):
    def point_assigner(points, gts, scale=4, pos_num=3):
        return op2p.point_assigner(points.as_subclass(torch.Tensor), gts.as_subclass(torch.Tensor), scale, pos_num)

    def cost_matrix(cls, pts, row_idx, gts, labels, w_cls, alpha, gamma, eps, w_dis, fx=1.0, fy=1.0, out=None):
        cfg = dict(cls_cost_weight=w_cls, focal_alpha=alpha, focal_gamma=gamma, dis_cost_weight=w_dis, dis_norm_with_img_wh=(fx, fy) != (1.0, 1.0),
                   dis_p=1)
        t = torch.Tensor
        return op2p.cost_matrix(pts.as_subclass(t), cls.as_subclass(t), gts.as_subclass(t), labels.as_subclass(t).long(), (fy, fx, 3), cfg)

    def hungarian(cost_flat, shapes, topk_k, out, out_offsets, row_idx=None, row_idx_offsets=None):
        (N, n), = shapes
        gi, _ = op2p.hungarian_v2_from_cost(cost_flat.as_subclass(torch.Tensor).view(N, n), torch.zeros(n, dtype=torch.long), topk_k)
        out[out_offsets[0]:out_offsets[0] + N] = gi
        return torch.zeros(1, dtype=torch.int32)

    def nms_boxes(boxes, scores, score_thr, iou_thr, max_per_img):
        t = torch.Tensor
        b, s = boxes.as_subclass(t)[0], scores.as_subclass(t)[0]
        full = torch.cat([s, s.new_zeros(s.shape[0], 1)], 1)
        dets, labels, keep, _ = op2p.multiclass_nms(b, full, score_thr, iou_thr, max_per_img)
        k = dets.shape[0]
        det = torch.zeros(1, max_per_img, 5); det[0, :k] = dets
        lab = torch.zeros(1, max_per_img, dtype=torch.int32); lab[0, :k] = labels.int()
        kp = torch.zeros(1, max_per_img, dtype=torch.int32)
        kp[0, :k] = keep.int()
        return torch.tensor([k], dtype=torch.int32), det, lab, kp, torch.tensor([int((s > score_thr).sum())], dtype=torch.int32)

    monkeypatch.setattr(ops, 'point_assigner', point_assigner)
    monkeypatch.setattr(ops, 'p2p_cost_matrix', cost_matrix)
    monkeypatch.setattr(ops, 'hungarian_v2_batch', hungarian)
    monkeypatch.setattr(ops, 'multiclass_nms_boxes', nms_boxes)


def test_point_assigner_mirror_reference_vectors(oracle_backend):
    A = assigners.PointAssigner()
    points = fc(torch.FloatTensor([[0, 0, 1], [10, 10, 1], [5, 5, 1], [32, 32, 1]]))
    gts = fc(torch.FloatTensor([[0, 0, 10, 9], [0, 10, 10, 19]]))
    r = A.assign(points, gts)                                                   # test_assigner.py:155-170
    assert r.gt_inds.tolist() == [1, 2, 1, 0] and r.labels is None and r.num_gts == 2
    r = A.assign(points, gts, gt_labels=fc(torch.LongTensor([7, 3])))
    assert r.labels.tolist() == [7, 3, 7, -1]
    r = A.assign(points, fc(torch.FloatTensor([])))                             # test_assigner.py:173-187
    assert r.gt_inds.tolist() == [0, 0, 0, 0]
    r = A.assign(fc(torch.FloatTensor([])), fc(torch.FloatTensor([])))          # test_assigner.py:190-196
    assert len(r.gt_inds) == 0
    with pytest.raises(RuntimeError, match='no CPU'):
        A.assign(torch.zeros(4, 3), torch.zeros(2, 4))


def test_hungarian_assigner_v2_mirror(oracle_backend):
    g = torch.Generator().manual_seed(0)
    N, n, C = 60, 5, 8
    pts, cls = torch.rand(N, 2, generator=g) * 100, torch.randn(N, C, generator=g)
    gts, labels = torch.rand(n, 2, generator=g) * 100, torch.randint(0, C, (n,), generator=g)
    registry.register_core()
    A = registry.build_assigner(dict(type='HungarianAssignerV2', cls_costs=dict(type='FocalLossCost', weight=2.0),
                                     reg_costs=dict(type='DisCostV2', weight=0.1, norm_with_img_wh=False), topk_k=5))
    r = A.assign(fc(pts), fc(cls), fc(gts), fc(labels), dict(img_shape=(100, 100, 3)))
    cfg = op2p.default_cfg(num_classes=C)
    cfg.update(cls_cost_weight=2.0, dis_cost_weight=0.1, dis_norm_with_img_wh=False)
    cost = op2p.cost_matrix(pts, cls, gts, labels, (100, 100, 3), cfg)
    gi, lb = op2p.hungarian_v2_from_cost(cost, labels, 5)
    assert torch.equal(r.gt_inds.as_subclass(torch.Tensor), gi) and torch.equal(r.labels.as_subclass(torch.Tensor), lb)
    assert int((r.gt_inds > 0).sum()) == 5 * n and r.num_gts == n
    r0 = A.assign(fc(pts), fc(cls), fc(torch.zeros(0, 2)), fc(torch.zeros(0, dtype=torch.long)), dict(img_shape=(100, 100, 3)))
    assert r0.gt_inds.tolist() == [0] * N and r0.labels.tolist() == [-1] * N
    S = registry.build_sampler(dict(type='PseudoSampler'))
    sr = S.sample(r, fc(pts), fc(gts))
    assert sr.pos_inds.tolist() == torch.nonzero(gi > 0).squeeze(1).tolist() and len(sr.neg_inds) == N - 5 * n
    assert torch.equal(sr.pos_gt_bboxes.as_subclass(torch.Tensor), gts[gi[gi > 0] - 1]) and sr.num_gts == n
    assert torch.equal(sr.pos_gt_labels.as_subclass(torch.Tensor), lb[gi > 0])
    with pytest.raises(NotImplementedError):
        assigners.HungarianAssignerV2(cls_costs=dict(type='ClassificationCost', weight=1.0), reg_costs=dict(type='DisCostV2'))
    with pytest.raises(NotImplementedError):
        assigners.HungarianAssignerV2()          # the reference's default (DETR) costs are not implemented: fail loudly


def test_multiclass_nms_mirror(oracle_backend):
    g = torch.Generator().manual_seed(1)
    n, C = 200, 4
    c = torch.rand(n, 2, generator=g) * 200
    boxes = torch.cat([c - 12, c + 12], 1)
    scores = torch.cat([torch.rand(n, C, generator=g) * (torch.rand(n, C, generator=g) > 0.6), torch.zeros(n, 1)], 1)
    dets, labels, keep = post_processing.multiclass_nms(fc(boxes), fc(scores), 0.05, dict(type='nms', iou_threshold=0.5), max_num=50, return_inds=True)
    rd, rl, rk = op2p.multiclass_nms(boxes, scores, 0.05, 0.5, 50)[:3]
    assert torch.equal(dets.as_subclass(torch.Tensor), rd) and torch.equal(labels.as_subclass(torch.Tensor), rl)
    assert torch.equal(keep.as_subclass(torch.Tensor), rk) and 0 < len(rk) <= 50
    d2, l2 = post_processing.multiclass_nms(fc(boxes), fc(scores), 0.05, dict(type='nms', iou_threshold=0.5), max_num=50)
    assert torch.equal(d2.as_subclass(torch.Tensor), rd)
    with pytest.raises(NotImplementedError):          # the kernel keeps at most 1024 detections (score_factors / max_num=-1: tests/test_gpu_p2p.py)
        post_processing.multiclass_nms(fc(boxes), fc(scores), 0.05, dict(type='nms', iou_threshold=0.5), max_num=2000)
    with pytest.raises(RuntimeError, match='no CPU'):
        post_processing.multiclass_nms(boxes, scores, 0.05, dict(type='nms', iou_threshold=0.5), max_num=50)


def test_core_registries():
    registry.register_core()
    assert {'HungarianAssignerV2', 'MaxIoUAssigner', 'PointAssigner'} <= set(registry.BBOX_ASSIGNERS.module_dict)
    a = registry.build_assigner(dict(type='MaxIoUAssigner', pos_iou_thr=0.7, neg_iou_thr=0.3, min_pos_iou=0.3, match_low_quality=True,
                                     ignore_iof_thr=-1))                       # faster_rcnn_r50_fpn_1x_TinyPerson640.py:56-62
    assert a.pos_iou_thr == 0.7
    ag = registry.build_anchor_generator(dict(type='AnchorGenerator', scales=[2], ratios=[0.5, 1.0, 2.0], strides=[4, 8, 16, 32, 64]))
    assert ag.num_base_anchors == [3] * 5 and torch.equal(ag.base_anchors[0], oa.base_anchors(4, [2], [0.5, 1.0, 2.0]))


@pytest.mark.parametrize('name', ['factors', 'class_specific', 'agnostic', 'unlimited'])
def test_multiclass_nms_mirror_options(oracle_backend, name, golden_dir):
    """the option glue of `post_processing.multiclass_nms` (score_factors folded into one ranking score, one kernel candidate per (box, class)
    for class-specific boxes / class_agnostic, max_num=-1) with the oracle standing in for the library, against the outputs recorded from
    the REAL reference function (tests/golden/multiclass_nms_options.npz); the GPU run of the same cases is in tests/test_gpu_p2p.py"""
    import os
    from tests.test_oracle_golden import NMS_OPTION_CASES
    c = NMS_OPTION_CASES[name]
    z = np.load(os.path.join(golden_dir, 'multiclass_nms_options.npz'))
    sf = fc(torch.from_numpy(z['factors'])) if c['sf'] else None
    d, l, k = post_processing.multiclass_nms(fc(torch.from_numpy(z[c['boxes']])), fc(torch.from_numpy(z['scores'])), 0.05, dict(c['cfg']),
                                            c['max_num'], score_factors=sf, return_inds=True)
    assert np.array_equal(k.as_subclass(torch.Tensor).numpy(), z[f'{name}_keep'])
    assert np.array_equal(l.as_subclass(torch.Tensor).numpy(), z[f'{name}_labels'])
    assert np.array_equal(d.as_subclass(torch.Tensor).numpy(), z[f'{name}_dets'])
