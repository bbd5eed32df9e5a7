"""a8 — grid-cell ("neighbour index") bags, GridCirclesPtFeatGenerator (reference cpr_head.py:296-350, 413-444).
CPU: the oracle restatement reproduces the fixtures recorded from the real reference (oracle/make_golden.py, grid_radius=...).
GPU: ptb_cpr_grid_bag (+ backward) and the CPRHead plugin configured with the grid generator against oracle and fixtures."""
import os

import numpy as np
import pytest
import torch

from oracle import cpr as ocpr, synth
from tests.helpers import assert_close, assert_mask_equal, flat_batch, oracle_cfg

CASES = [('lite', 1234, 3), ('mid', 77, 2)]


def _cfg(d, radius):
    cfg = oracle_cfg(d)
    cfg.update(pos_generator='grid_circles', pos_radius=radius)
    return cfg


def _gold_chosens(gold):
    shp = gold['chosens_shape']
    return np.unpackbits(gold['chosens'])[:int(np.prod(shp))].reshape(shp).astype(bool)


@pytest.mark.parametrize('name,seed,radius', CASES)
def test_oracle_grid_bags_match_reference_golden(golden_dir, name, seed, radius):
    inp = synth.cpr_inputs(name, seed)
    cfg = _cfg(inp['cfgd'], radius)
    gold = np.load(os.path.join(golden_dir, f'cpr_{name}_grid.npz'))
    assert int(gold['seed']) == seed and int(gold['grid_radius']) == radius
    gt_r = [p.reshape(len(l), -1, 2) for p, l in zip(ocpr.pseudo_bbox_to_center(inp['gt_bboxes']), inp['gt_labels'])]
    ex = ocpr.extract(inp['cls_feat'], gt_r, inp['gt_labels'], inp['img_metas'], cfg)
    assert ex['pos_pts'].shape[2] == 2 * (2 * radius) ** 2 + 2          # max_pos_num + 2*num_refine slots
    assert np.array_equal(ex['pos_valid'].numpy(), gold['pos_valid'])
    assert np.array_equal(ex['pos_pts'].numpy(), gold['pos_pts'])
    assert np.array_equal(ex['pos_feats'].flatten()[::1009].numpy(), gold['pos_feats_sub'])
    ch = []
    for b, m in enumerate(inp['img_metas']):
        ch.append(ocpr.grid_circles_bag(inp['cls_feat'][b:b + 1], gt_r[b], *m['pad_shape'][:2], cfg['stride'], radius,
                                        keep_feats=False)[3].flatten(1))
    assert np.array_equal(torch.cat(ch).numpy(), _gold_chosens(gold))
    losses = ocpr.cpr_loss(inp['cls_feat'], inp['weights'], inp['gt_bboxes'], inp['gt_labels'], inp['img_metas'], cfg)
    for k in ('gt_loss', 'pos_loss', 'neg_loss', 'bag_acc'):
        np.testing.assert_allclose(losses[k].reshape(-1).numpy(), gold['loss_' + k], rtol=1e-6)
    res, ra = ocpr.cpr_get_bboxes(inp['cls_feat'], inp['weights'], inp['gt_bboxes'], inp['gt_labels'], inp['gt_anns_id'],
                                  inp['img_metas'], cfg, return_all=True)
    assert np.array_equal(torch.cat([r[0] for r in res]).numpy(), gold['det'])
    for key in ('not_refine', 'chosen', 'merge_valid'):
        assert np.array_equal(torch.cat([r[key] for r in ra['refine']]).numpy(), gold[key]), key


def test_oracle_grid_bag_overflow_raises_like_the_reference():
    """more chosen cells than max_pos_num + num_refine slots: the reference's slice assignment fails (cpr_head.py:334)"""
    feat = torch.zeros(1, 4, 16, 16)
    with pytest.raises(RuntimeError):
        ocpr.grid_circles_bag(feat, torch.tensor([[[64., 64.]]]), 128, 128, 8, 3, max_pos_num=5)


# ------------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def ops():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from pointtinybenchmark_b200 import ops
    return ops


@pytest.mark.gpu
@pytest.mark.parametrize('name,seed,radius', CASES)
def test_grid_bag_kernel_vs_oracle_and_golden(ops, golden_dir, name, seed, radius):
    dev = torch.device('cuda:0')
    inp = synth.cpr_inputs(name, seed)
    d = inp['cfgd']
    cfg = _cfg(d, radius)
    gold = np.load(os.path.join(golden_dir, f'cpr_{name}_grid.npz'))
    fb, lens = flat_batch(inp, dev)
    fmap = ops.to_nhwc(inp['cls_feat'].to(dev))
    B, H, W, C = fmap.shape
    feats, pts, valid, cell = ops.grid_bag(fmap, fb['centers'], fb['bag_img'], d['stride'], radius)
    gt_r = [p.reshape(len(l), -1, 2) for p, l in zip(ocpr.pseudo_bbox_to_center(inp['gt_bboxes']), inp['gt_labels'])]
    ex = ocpr.extract(inp['cls_feat'], gt_r, inp['gt_labels'], inp['img_metas'], cfg)
    # neighbour index / validity / coordinates: bit exact
    assert torch.equal(pts.cpu(), ex['pos_pts'][:, 0]), 'bag points'
    assert_mask_equal(valid, ex['pos_valid'][:, 0, :, 0], 'bag valid')
    assert np.array_equal(pts.cpu().numpy(), gold['pos_pts'][:, 0])
    assert np.array_equal(valid.cpu().numpy(), gold['pos_valid'][:, 0, :, 0])
    ch = torch.zeros(cell.shape[0], H * W, dtype=torch.bool)
    cc = cell.cpu().long()
    rows = torch.arange(cc.shape[0])[:, None].expand_as(cc)
    ch[rows[cc >= 0], cc[cc >= 0]] = True
    assert np.array_equal(ch.numpy(), _gold_chosens(gold)), 'chosens (neighbour mask) vs the reference'
    assert bool((cell[:, -1] == -2).all()) and bool(((cell[:, :-1] >= 0) == valid[:, :-1]).all())
    body, bv = cc[:, :-1], valid[:, :-1].cpu()
    assert bool((bv[:, :-1] | ~bv[:, 1:]).all()), 'filled slots form a prefix'
    both = bv[:, :-1] & bv[:, 1:]
    assert bool((body.diff(dim=1)[both] > 0).all()), 'row-major order'
    # cell vectors are exact copies, the centre is the bit-exact bilinear sample
    assert torch.equal(feats.cpu(), ex['pos_feats'][:, 0]), 'bag features'
    assert np.array_equal(feats.cpu().flatten()[::1009].numpy(), gold['pos_feats_sub'])
    # backward == autograd through the oracle's indexing
    g = torch.Generator().manual_seed(5)
    go = torch.randn(feats.shape, generator=g)
    gm = ops.grid_bag_bwd(go.to(dev), (B, H, W, C), fb['centers'], fb['bag_img'], cell, d['stride'])
    fo = inp['cls_feat'].clone().requires_grad_(True)
    ex2 = ocpr.extract(fo, gt_r, inp['gt_labels'], inp['img_metas'], cfg)
    (ex2['pos_feats'][:, 0] * go).sum().backward()
    assert_close(gm.permute(0, 3, 1, 2), fo.grad, 1e-5, 'grid bag backward')


@pytest.mark.gpu
def test_grid_bag_edges(ops):
    dev = torch.device('cuda:0')
    fmap = torch.randn(2, 12, 20, 8, device=dev)
    centers = torch.tensor([[-30., -30.], [4., 4.], [159.9, 95.9], [500., 40.], [80., 48.]], device=dev)
    bag_img = torch.tensor([0, 0, 1, 1, 1], dtype=torch.int32, device=dev)
    f, p, v, c = ops.grid_bag(fmap, centers, bag_img, 8, 2)
    assert f.shape == (5, 34, 8) and int(v[0, :-1].sum()) == 0 and int(v[3, :-1].sum()) == 0     # far outside: only the centre
    assert bool(v[:, -1].all()) and int(v[4, :-1].sum()) == 12                                  # (+-4,+-4), (+-4,+-12), (+-12,+-4)
    o = ocpr.grid_circles_bag(fmap[1:2].permute(0, 3, 1, 2).cpu(), centers[2:].cpu()[:, None], 96, 160, 8, 2)
    assert torch.equal(p[2:].cpu(), o[0][:, 0]) and torch.equal(f[2:].cpu(), o[2][:, 0])
    with pytest.raises(RuntimeError, match='max_pos_num'):
        ops.grid_bag(fmap, centers, bag_img, 8, 2, max_pos_num=5)
    with pytest.raises(TypeError):
        ops.grid_bag(fmap, centers, bag_img, 8, 2.5)
    e = ops.grid_bag(fmap, centers[:0], bag_img[:0], 8, 2)
    assert e[0].shape == (0, 34, 8)


@pytest.mark.gpu
@pytest.mark.parametrize('name,seed,radius', CASES)
def test_cpr_head_with_grid_generator(golden_dir, name, seed, radius):
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from pointtinybenchmark_b200 import cpr_head  # noqa: F401
    from pointtinybenchmark_b200.registry import build_head
    from tests.test_gpu_cpr_head import head_cfg
    dev = torch.device('cuda:0')
    inp = synth.cpr_inputs(name, seed)
    cfg = _cfg(inp['cfgd'], radius)
    gold = np.load(os.path.join(golden_dir, f'cpr_{name}_grid.npz'))
    hc = head_cfg(inp['cfgd'])
    for ex in ('train_pts_extractor', 'refine_pts_extractor'):
        hc[ex]['pos_generator'] = dict(type='GridCirclesPtFeatGenerator', radius=radius)
    head = build_head(hc).cuda()
    sd = head.state_dict()
    sd.update(inp['weights'])
    head.load_state_dict(sd, strict=True)
    gtb, gtl, aid = ([b.to(dev) for b in inp['gt_bboxes']], [l.to(dev) for l in inp['gt_labels']],
                     [a.to(dev) for a in inp['gt_anns_id']])
    feat = inp['cls_feat'].to(dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    losses = head.loss([feat], [feat], gtb, gtl, inp['img_metas'])
    sum(v for k, v in losses.items() if 'loss' in k).backward()
    fo = inp['cls_feat'].clone().requires_grad_(True)
    wo = {k: v.clone().requires_grad_(True) for k, v in inp['weights'].items()}
    ol = ocpr.cpr_loss(fo, wo, inp['gt_bboxes'], inp['gt_labels'], inp['img_metas'], cfg)
    sum(v for k, v in ol.items() if 'loss' in k).backward()
    for k in ('gt_loss', 'pos_loss', 'neg_loss', 'bag_acc'):
        assert_close(losses[k].reshape(-1), ol[k].detach().reshape(-1), 1e-4, k)
        assert_close(losses[k].reshape(-1), torch.from_numpy(gold['loss_' + k]), 1e-4, k + ' vs golden')
    assert_close(feat.grad, fo.grad, 2e-4, 'd loss / d feature map')
    assert_close(head.cls_out.weight.grad, wo['cls_out.weight'].grad, 2e-4, 'd/d cls_out.weight')
    assert_close(head.ins_out.weight.grad, wo['ins_out.weight'].grad, 2e-4, 'd/d ins_out.weight')
    head.eval()
    with torch.no_grad():
        res, nr = head.get_bboxes([feat.detach()], [feat.detach()], inp['img_metas'], gt_bboxes=gtb, gt_labels=gtl,
                                  gt_anns_id=aid, cascade_out_fmt=True)
    assert_mask_equal(torch.cat(nr), torch.from_numpy(gold['not_refine']), 'not_refine vs golden')
    assert_close(torch.cat([r[0] for r in res]), torch.from_numpy(gold['det']), 1e-4, 'det vs golden')
