"""GPU: the assigner / NMS plugin mirrors (pointtinybenchmark_b200/{assigners,post_processing}.py) over the real library, against the
oracle and the reference's own vectors.  (File name sorts last: these wrappers were added after the round's last GPU session; the
kernels underneath are covered by the other -m gpu tests.)"""
import pytest
import torch

from oracle import p2p as op2p


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda:0')


@pytest.mark.gpu
def test_point_assigner_mirror_gpu(dev):
    from pointtinybenchmark_b200.assigners import PointAssigner
    A = PointAssigner()
    points = torch.FloatTensor([[0, 0, 1], [10, 10, 1], [5, 5, 1], [32, 32, 1]]).to(dev)
    gts = torch.FloatTensor([[0, 0, 10, 9], [0, 10, 10, 19]]).to(dev)
    r = A.assign(points, gts, gt_labels=torch.LongTensor([7, 3]).to(dev))      # test_assigner.py:155-170
    assert r.gt_inds.tolist() == [1, 2, 1, 0] and r.labels.tolist() == [7, 3, 7, -1]
    assert A.assign(points, torch.FloatTensor([]).to(dev)).gt_inds.tolist() == [0, 0, 0, 0]
    assert len(A.assign(torch.FloatTensor([]).to(dev), torch.FloatTensor([]).to(dev)).gt_inds) == 0
    g = torch.Generator().manual_seed(3)
    pts = torch.cat([torch.rand(500, 2, generator=g) * 256, torch.tensor([8., 16., 32.])[torch.randint(0, 3, (500, 1), generator=g)]], 1)
    c = torch.rand(9, 2, generator=g) * 256
    wh = torch.rand(9, 2, generator=g) * 60 + 4
    gb = torch.cat([c - wh / 2, c + wh / 2], 1)
    assert torch.equal(A.assign(pts.to(dev), gb.to(dev)).gt_inds.cpu(), op2p.point_assigner(pts, gb))


@pytest.mark.gpu
def test_hungarian_assigner_v2_mirror_gpu(dev):
    from pointtinybenchmark_b200.assigners import HungarianAssignerV2, PseudoSampler
    g = torch.Generator().manual_seed(0)
    N, n, C = 700, 9, 8
    pts, cls = torch.rand(N, 2, generator=g) * 100, torch.randn(N, C, generator=g)
    gts, labels = torch.rand(n, 2, generator=g) * 100, torch.randint(0, C, (n,), generator=g)
    A = HungarianAssignerV2(cls_costs=dict(type='FocalLossCost', weight=2.0), reg_costs=dict(type='DisCostV2', weight=0.1, norm_with_img_wh=False),
                            topk_k=5)
    r = A.assign(pts.to(dev), cls.to(dev), gts.to(dev), labels.to(dev), dict(img_shape=(100, 100, 3)))
    cfg = op2p.default_cfg(num_classes=C)
    cfg.update(cls_cost_weight=2.0, dis_cost_weight=0.1, dis_norm_with_img_wh=False)
    # the assignment for OUR cost matrix must be scipy's (bit-exact); the cost itself is compared with the oracle at 1e-4 elsewhere
    from pointtinybenchmark_b200 import ops
    cost = ops.p2p_cost_matrix(cls.to(dev), pts.to(dev), None, gts.to(dev), labels.int().to(dev), 2.0, 0.25, 2.0, 1e-12, 0.1, 1.0, 1.0).cpu()
    gi, lb = op2p.hungarian_v2_from_cost(cost, labels, 5)
    assert torch.equal(r.gt_inds.cpu(), gi) and torch.equal(r.labels.cpu(), lb) and int((gi > 0).sum()) == 5 * n
    sr = PseudoSampler().sample(r, pts.to(dev), gts.to(dev))
    assert torch.equal(sr.pos_gt_bboxes.cpu(), gts[gi[gi > 0] - 1]) and len(sr.neg_inds) == N - 5 * n


@pytest.mark.gpu
def test_multiclass_nms_mirror_gpu(dev):
    from pointtinybenchmark_b200.post_processing import multiclass_nms
    g = torch.Generator().manual_seed(1)
    n, C = 300, 4
    c = torch.rand(n, 2, generator=g) * 200
    boxes = torch.cat([c - 12, c + 12], 1)
    scores = torch.cat([torch.rand(n, C, generator=g) * (torch.rand(n, C, generator=g) > 0.6), torch.zeros(n, 1)], 1)
    dets, labels, keep = multiclass_nms(boxes.to(dev), scores.to(dev), 0.05, dict(type='nms', iou_threshold=0.5), max_num=50, return_inds=True)
    rd, rl, rk = op2p.multiclass_nms(boxes, scores, 0.05, 0.5, 50)[:3]
    assert torch.equal(keep.cpu(), rk) and torch.equal(labels.cpu(), rl) and torch.equal(dets.cpu(), rd)
