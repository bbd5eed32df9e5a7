"""RPN proposal path for dense anchors: AnchorGenerator grid anchors, DeltaXYWHBBoxCoder.decode, per-level
top-k and batched NMS (anchor_head.py:551-590, rpn_head.py:78-186).
CPU: the oracle restatement reproduces the fixture recorded from the REAL reference RPNHead.get_bboxes / AnchorGenerator
(oracle/make_golden.py::golden_rpn); the host mirror's anchors equal the reference's.
GPU: ptb_rpn_proposals — top-k anchor indices and NMS decisions bit-exact, boxes within 1e-4 (the decode contains an exp)."""
import os

import numpy as np
import pytest
import torch

from oracle import anchors as oa
from oracle import p2p as op2p

CASES = [('a', 3, (512, 640), 1000, 1000), ('b', 4, (256, 320), 300, 100), ('c', 5, (64, 96), 1000, 50)]


def test_oracle_matches_reference_golden(golden_dir):
    gold = np.load(os.path.join(golden_dir, 'rpn_proposals.npz'))
    for name, seed, size, nms_pre, max_per_img in CASES:
        cfg = dict(oa.RPN_CFG, nms_pre=nms_pre, max_per_img=max_per_img)
        cls, box, shapes = oa.synth_rpn_inputs(seed, size=size)
        dets, al = oa.rpn_proposals(cls, box, shapes, cfg, return_all=True)
        for b in range(len(dets)):
            assert np.array_equal(dets[b].numpy(), gold[f'{name}_dets{b}']), (name, b)
            assert np.array_equal(al['per_image'][b]['keep_pos'].numpy(), gold[f'{name}_keep_pos{b}'])
        assert np.array_equal(al['cand_idx'].numpy(), gold[f'{name}_cand_idx'])
        assert np.array_equal(al['cand_boxes'].reshape(-1)[::5].numpy(), gold[f'{name}_cand_boxes_sub'])


def test_delta2bbox_docstring_vector():
    """the reference's own known-answer example (delta_xywh_bbox_coder.py:190-203)"""
    rois = torch.Tensor([[0., 0., 1., 1.], [0., 0., 1., 1.], [0., 0., 1., 1.], [5., 5., 5., 5.]])
    deltas = torch.Tensor([[0., 0., 0., 0.], [1., 1., 1., 1.], [0., 0., 2., -1.], [0.7, -1.9, -0.5, 0.3]])
    want = torch.Tensor([[0.0000, 0.0000, 1.0000, 1.0000], [0.1409, 0.1409, 2.8591, 2.8591], [0.0000, 0.3161, 4.1945, 0.6839],
                         [5.0000, 5.0000, 5.0000, 5.0000]])
    assert torch.allclose(oa.delta2bbox(rois, deltas, max_shape=(32, 32, 3)), want, atol=1e-4)


def test_host_anchor_generator_matches_reference(golden_dir):
    from pointtinybenchmark_b200.rpn import AnchorGenerator
    c = oa.RPN_CFG
    ag = AnchorGenerator(scales=c['scales'], ratios=c['ratios'], strides=c['strides'])
    assert ag.num_base_anchors == [3] * 5 and ag.num_levels == 5
    for l, s in enumerate(c['strides']):
        assert torch.equal(ag.base_anchors[l], oa.base_anchors(s, c['scales'], c['ratios']))           # oracle is pinned to the reference
        assert torch.equal(ag.single_level_grid_anchors(ag.base_anchors[l], (7, 5), (s, s), device='cpu'),
                           oa.grid_anchors(oa.base_anchors(s, c['scales'], c['ratios']), (7, 5), (s, s)))
    gold = np.load(os.path.join(golden_dir, 'rpn_proposals.npz'))
    vf = ag.valid_flags([(16, 20), (8, 10), (4, 5), (2, 3), (1, 2)], (60, 75, 3), device='cpu')
    assert np.array_equal(torch.cat(vf).numpy(), gold['valid_flags'])
    # octave-style ctor (RetinaNet configs) builds too
    ag2 = AnchorGenerator(strides=[8, 16], ratios=[0.5, 1.0, 2.0], octave_base_scale=4, scales_per_octave=3)
    assert ag2.num_base_anchors == [9, 9]
    from pointtinybenchmark_b200.rpn import RPNProposals
    with pytest.raises(RuntimeError, match='no CPU'):
        cls, box, shapes = oa.synth_rpn_inputs(1, size=(64, 96))
        RPNProposals(dict(type='AnchorGenerator', scales=c['scales'], ratios=c['ratios'], strides=c['strides']),
                     test_cfg=dict(nms_pre=100, max_per_img=10, nms=dict(type='nms', iou_threshold=0.7), min_bbox_size=0)
                     ).get_bboxes(cls, box, [dict(img_shape=s) for s in shapes])


def _close(a, b, tol=1e-4):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.all(np.abs(a - b) <= tol * np.maximum(np.abs(b), 1.0))


@pytest.mark.gpu
@pytest.mark.parametrize('case', CASES + [('d', 6, (512, 640), 2000, 1000), ('e', 7, (96, 128), -1, 64)], ids=lambda c: c[0])
def test_rpn_proposals_kernel(case):
    from pointtinybenchmark_b200 import ops
    from pointtinybenchmark_b200.rpn import AnchorGenerator, RPNProposals
    name, seed, size, nms_pre, max_per_img = case
    dev = torch.device('cuda:0')
    c = oa.RPN_CFG
    cfg = dict(c, nms_pre=nms_pre, max_per_img=max_per_img)
    strides = c['strides'] if name != 'e' else [8, 16, 32, 64, 128]        # case e: nms_pre=-1 keeps every anchor (<= 4096 per level)
    cfg['strides'] = strides
    cls, box, shapes = oa.synth_rpn_inputs(seed, size=size, strides=strides)
    odets, al = oa.rpn_proposals(cls, box, shapes, cfg, return_all=True)
    ag = AnchorGenerator(scales=c['scales'], ratios=c['ratios'], strides=strides)
    img_hw = torch.tensor([[s[0], s[1]] for s in shapes], dtype=torch.int32, device=dev)
    cnt, det, lvl, ex = ops.rpn_proposals([t.to(dev) for t in cls], [t.to(dev) for t in box], torch.stack(ag.base_anchors).to(dev),
                                          ag.strides, img_hw, c['means'], c['stds'], 16 / 1000, nms_pre, c['min_bbox_size'],
                                          c['iou_threshold'], max_per_img, want_candidates=True)
    torch.cuda.synchronize()
    B = len(shapes)
    # 1. per-level top-k anchor indices: identical up to the order INSIDE groups of exactly equal reference scores (two logits
    #    whose fp32 sigmoids collide: the reference's torch.sort(descending=True) gives no order contract there, and a sigmoid that
    #    differs by 1 ulp separates them);  scores 1e-6, boxes 1e-4 after aligning the two lists by anchor index
    seg = [min(nms_pre, t.shape[1] * t.shape[2] * t.shape[3]) if nms_pre > 0 else t.shape[1] * t.shape[2] * t.shape[3] for t in cls]
    ids = torch.cat([torch.full((n,), l, dtype=torch.long) for l, n in enumerate(seg)])
    off = np.concatenate([[0], np.cumsum(seg)])
    ex_idx, or_idx, or_sc = ex['cand_idx'].cpu().numpy(), al['cand_idx'].numpy(), al['cand_scores'].numpy()
    perm = np.zeros_like(ex_idx)                                  # our position -> the oracle's position of the same anchor
    n_swapped = 0
    for b in range(B):
        for l in range(len(seg)):
            e, o, sc_o = ex_idx[b, off[l]:off[l + 1]], or_idx[b, off[l]:off[l + 1]], or_sc[b, off[l]:off[l + 1]]
            where = {int(a): r for r, a in enumerate(o)}
            assert sorted(e.tolist()) == sorted(o.tolist()), f'level {l}: selected anchor set'
            pr = np.array([where[int(a)] for a in e])
            bad = np.nonzero(pr != np.arange(len(e)))[0]
            assert all(sc_o[pr[r]] == sc_o[r] for r in bad), 'order differs outside a group of equal reference scores'
            n_swapped += len(bad)
            perm[b, off[l]:off[l + 1]] = pr + off[l]
    assert n_swapped <= 8
    bi = np.arange(B)[:, None]
    assert _close(ex['cand_score'].cpu().numpy(), or_sc[bi, perm], 1e-6)
    assert _close(ex['cand_box'].cpu().numpy(), al['cand_boxes'].numpy()[bi, perm], 1e-4)
    for b in range(B):
        n = int(cnt[b])
        # 2. NMS decisions: the reference algorithm (oracle batched_nms) on OUR candidate list must give exactly our output
        p, sc = ex['cand_box'][b].cpu(), ex['cand_score'][b].cpu()
        v = torch.nonzero(((p[:, 2] - p[:, 0]) > c['min_bbox_size']) & ((p[:, 3] - p[:, 1]) > c['min_bbox_size'])).squeeze(1)
        dets, keep = op2p.batched_nms(p[v], sc[v], ids[v], c['iou_threshold'])
        dets, keep = dets[:max_per_img], v[keep][:max_per_img]
        assert n == dets.shape[0], (name, b, n, dets.shape)
        assert torch.equal(ex['pos'][b, :n].cpu().long(), keep), 'NMS keep positions'
        assert torch.equal(det[b, :n].cpu(), dets), 'dets'
        assert torch.equal(lvl[b, :n].cpu().long(), ids[keep])
        # 3. end to end against the reference-pinned oracle: the same anchors survive, values within 1e-4, scores non-increasing
        assert n == odets[b].shape[0]
        ours = {(int(ids[q]), int(ex_idx[b, q])): r for r, q in enumerate(keep.tolist())}
        kp = al['per_image'][b]['keep_pos'].tolist()
        theirs = [(int(ids[q]), int(or_idx[b, q])) for q in kp]
        assert set(ours) == set(theirs), 'kept anchors vs oracle (would differ only if an IoU sits within 1e-6 of the threshold)'
        rows = [ours[k] for k in theirs]
        assert _close(det[b, :n].cpu().numpy()[rows], odets[b].numpy(), 1e-4)
        assert bool((det[b, 1:n, 4] <= det[b, :n - 1, 4]).all())
    # host mirror: RPNHead.get_bboxes signature
    rp = RPNProposals(dict(type='AnchorGenerator', scales=c['scales'], ratios=c['ratios'], strides=strides),
                      dict(type='DeltaXYWHBBoxCoder', target_means=list(c['means']), target_stds=list(c['stds'])),
                      test_cfg=dict(nms_pre=nms_pre, max_per_img=max_per_img, nms=dict(type='nms', iou_threshold=c['iou_threshold']),
                                    min_bbox_size=c['min_bbox_size']))
    res = rp.get_bboxes([t.to(dev) for t in cls], [t.to(dev) for t in box], [dict(img_shape=s, scale_factor=1.0) for s in shapes])
    assert len(res) == B and all(torch.equal(res[b], det[b, :int(cnt[b])]) for b in range(B))


@pytest.mark.gpu
def test_rpn_proposals_min_bbox_size_and_small_batch():
    from pointtinybenchmark_b200 import ops
    from pointtinybenchmark_b200.rpn import AnchorGenerator
    dev = torch.device('cuda:0')
    c = oa.RPN_CFG
    cfg = dict(c, nms_pre=200, max_per_img=30, min_bbox_size=6)
    cls, box, shapes = oa.synth_rpn_inputs(9, B=1, size=(128, 160))
    odets, al = oa.rpn_proposals(cls, box, shapes, cfg, return_all=True)
    ag = AnchorGenerator(scales=c['scales'], ratios=c['ratios'], strides=c['strides'])
    img_hw = torch.tensor([[s[0], s[1]] for s in shapes], dtype=torch.int32, device=dev)
    cnt, det, lvl, ex = ops.rpn_proposals([t.to(dev) for t in cls], [t.to(dev) for t in box], torch.stack(ag.base_anchors).to(dev), ag.strides,
                                          img_hw, c['means'], c['stds'], 16 / 1000, 200, 6, c['iou_threshold'], 30, want_candidates=True)
    n = int(cnt[0])
    assert n == odets[0].shape[0] and torch.equal(ex['pos'][0, :n].cpu().long(), al['per_image'][0]['keep_pos'])
    assert _close(det[0, :n].cpu().numpy(), odets[0].numpy(), 1e-4)
    w = det[0, :n, 2] - det[0, :n, 0]
    assert bool((w > 6).all())


def test_level_decomposition_of_the_joint_nms_is_exact():
    """design claim of rpn.cu: after batched_nms's level offset the levels are disjoint (all boxes are clipped to >= 0), so the joint
    greedy NMS equals per-level NMS on the SAME offset coordinates merged by (score desc, position asc) — checked on the oracle."""
    for name, seed, size, nms_pre, max_per_img in CASES:
        cfg = dict(oa.RPN_CFG, nms_pre=nms_pre, max_per_img=max_per_img)
        cls, box, shapes = oa.synth_rpn_inputs(seed, size=size)
        _, al = oa.rpn_proposals(cls, box, shapes, cfg, return_all=True)
        seg = [min(nms_pre, t.shape[1] * t.shape[2] * t.shape[3]) for t in cls]
        ids = torch.cat([torch.full((n,), l, dtype=torch.long) for l, n in enumerate(seg)])
        for b in range(len(shapes)):
            p, sc = al['cand_boxes'][b], al['cand_scores'][b]
            v = torch.nonzero(((p[:, 2] - p[:, 0]) > 0) & ((p[:, 3] - p[:, 1]) > 0)).squeeze(1)
            pv, sv, iv = p[v], sc[v], ids[v]
            off = iv.to(pv) * (pv.max() + 1)
            kept = []
            for l in range(len(seg)):
                m = torch.nonzero(iv == l).squeeze(1)
                if len(m):
                    k = op2p.nms(pv[m] + off[m, None], sv[m], cfg['iou_threshold'])
                    k = k[1] if isinstance(k, tuple) else k
                    kept.append(m[k][:max_per_img])
            kept = torch.cat(kept)
            order = sorted(kept.tolist(), key=lambda q: (-float(sv[q]), q))[:max_per_img]
            assert v[torch.tensor(order)].tolist() == al['per_image'][b]['keep_pos'].tolist(), (name, b)


@pytest.mark.gpu
@pytest.mark.parametrize('case', CASES, ids=lambda c: c[0])
def test_bitmask_nms_equals_the_serial_kernel(case, monkeypatch):
    """Round 2: `rpn_nms_level_bitmask_kernel` + `rpn_nms_merge_rank_kernel` (every level <= 1024 candidates) against the round-1 serial
    greedy kernel (PTB_RPN_NMS=serial) on the same inputs: every output identical, bit for bit (same greedy order, same IoU predicate)."""
    from pointtinybenchmark_b200 import ops
    from pointtinybenchmark_b200.rpn import AnchorGenerator
    name, seed, size, nms_pre, max_per_img = case
    dev = torch.device('cuda:0')
    c = oa.RPN_CFG
    cls, box, shapes = oa.synth_rpn_inputs(seed + 100, size=size, strides=c['strides'])
    ag = AnchorGenerator(scales=c['scales'], ratios=c['ratios'], strides=c['strides'])
    img_hw = torch.tensor([[s[0], s[1]] for s in shapes], dtype=torch.int32, device=dev)

    def run():
        out = ops.rpn_proposals([t.to(dev) for t in cls], [t.to(dev) for t in box], torch.stack(ag.base_anchors).to(dev), ag.strides, img_hw,
                                c['means'], c['stds'], 16 / 1000, nms_pre, c['min_bbox_size'], c['iou_threshold'], max_per_img,
                                want_candidates=True)
        torch.cuda.synchronize()
        return out

    monkeypatch.delenv('PTB_RPN_NMS', raising=False)
    cnt_a, det_a, lvl_a, ex_a = run()
    monkeypatch.setenv('PTB_RPN_NMS', 'serial')
    cnt_b, det_b, lvl_b, ex_b = run()
    assert torch.equal(cnt_a, cnt_b) and int(cnt_a.min()) > 0
    for b in range(cnt_a.numel()):
        n = int(cnt_a[b])
        assert torch.equal(det_a[b, :n], det_b[b, :n]) and torch.equal(lvl_a[b, :n], lvl_b[b, :n])
        if 'pos' in ex_a:
            assert torch.equal(ex_a['pos'][b, :n], ex_b['pos'][b, :n])
