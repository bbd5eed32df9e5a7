"""shared helpers for the parity tests (test infrastructure)."""
import numpy as np
import torch

from oracle import cpr as ocpr


def scale_rel_err(a, b):
    """max |a-b| / max(|b|, rms(b))  — the scale-relative metric of SURVEY.md §7.2 (tolerance 1e-4)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    if b.numel() == 0:
        return 0.0
    rms = float(b.pow(2).mean().sqrt())
    return float(((a - b).abs() / torch.clamp(b.abs(), min=max(rms, 1e-30))).max())


def assert_close(a, b, tol=1e-4, what=''):
    e = scale_rel_err(a, b)
    assert e <= tol, f'{what}: scale-relative error {e:.3e} > {tol}'
    return e


def assert_mask_equal(a, b, what=''):
    a, b = a.detach().cpu().bool(), b.detach().cpu().bool()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    nbad = int((a != b).sum())
    assert nbad == 0, f'{what}: {nbad} / {a.numel()} mask elements differ'


def flat_batch(inp, device):
    """concatenate the per-image GT lists of a synth batch into the CSR device tensors the ops take."""
    gtb, gtl, metas = inp['gt_bboxes'], inp['gt_labels'], inp['img_metas']
    centers = torch.cat([(b[:, :2] + b[:, 2:]) / 2 for b in gtb]).contiguous()
    labels = torch.cat(gtl).int()
    lens = [len(l) for l in gtl]
    bag_img = torch.cat([torch.full((n,), i, dtype=torch.int32) for i, n in enumerate(lens)])
    img_ptr = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32)
    pad_hw = torch.tensor([m['pad_shape'][:2] for m in metas], dtype=torch.int32)
    img_hw = torch.tensor([m['img_shape'][:2] for m in metas], dtype=torch.int32)
    d = dict(centers=centers, labels=labels, bag_img=bag_img, img_ptr=img_ptr, pad_hw=pad_hw, img_hw=img_hw)
    return {k: v.to(device) for k, v in d.items()}, lens


def oracle_cfg(d):
    return ocpr.default_cfg(num_classes=d['num_classes'], in_channels=d['C'], feat_channels=d['C'], stride=d['stride'],
                            pos_radius=d['radius'], neg_radius=d['radius'])


# ------------------------------------------------------------------------------------------------------------------------
# decision-margin harness: integer outputs of the refine step are thresholded / arg-max'd floats.  The product path
# computes the class logits in a different (mathematically identical) order than the reference — Linear before bilinear sampling,
# tensor-core towers — so its probabilities differ from the oracle's by `delta` (<= 1e-4 scale-relative by the float tolerance).  A
# comparison whose float64 margin exceeds the bound cannot legitimately flip: there the masks must be BIT-EQUAL; the (tiny) set of
# comparisons inside the bound is counted and reported, not hidden.
# ------------------------------------------------------------------------------------------------------------------------
def refine_decision_margins(bag_prob, labels, cfg):
    """bag_prob (G,K,C) oracle probabilities (centre sample last), labels (G,) -> float64 (G,K): the smallest margin of the
    probability-dependent comparisons of PointRefiner.refine_single (cpr_head.py:745-756 classify filter arg-max, :811-813 merge_th and
    gt_alpha thresholds).  The coordinate-only masks (valid, nearest GT, inside image) do not depend on the logits."""
    p64 = bag_prob.double()
    G, K, C = p64.shape
    gi = torch.arange(G)
    p = p64[gi, :, labels]                                              # (G,K) probability of the GT's own class
    pg = p[:, K - 1:K]
    m = torch.minimum((p - float(cfg['merge_th'])).abs(), (p - pg * float(cfg['gt_alpha'])).abs())
    if cfg['classify_filter'] and C > 1:
        other = p64.clone()
        other[gi, :, labels] = -1.0
        m = torch.minimum(m, (p - other.max(dim=-1)[0]).abs())
    return m


def check_refine_against_oracle(got, ora, bag_prob, labels, cfg, bound, what=''):
    """got = (pts (G,2), scores (G,), not_refine (G,), chosen (G,K)) of the CUDA path; ora = dict of the oracle's refine_single outputs
    concatenated over images; asserts bit-equality of every decision whose margin exceeds `bound`, 1e-4 on the floats of every GT
    whose decisions all agree, and returns the counts that were inside the bound."""
    pts, sc, nr, ch = [t.detach().cpu() for t in got]
    G, K = ch.shape
    m = refine_decision_margins(bag_prob, labels, cfg)
    exact = ora['mask_inside'].bool()                                   # masks that depend on coordinates only
    for key in ('mask_nearest', 'mask_valid'):
        if key in ora:
            exact = exact & ora[key].bool()
    decisive = (m > bound) | ~exact                                     # a sample outside the exact masks is never chosen, whatever p is
    flips = ch.bool() != ora['chosen'].bool()
    bad = flips & decisive
    assert int(bad.sum()) == 0, (f'{what}: {int(bad.sum())} chosen-mask flips with margin > {bound:.1e} '
                                 f'(smallest offending margin {float(m[bad].min()):.3e})')
    rows_equal = ~flips.any(dim=1)
    sm = (ora['refine_scores'].double() - float(cfg['refine_th'])).abs()
    # `not_refine = score < refine_th` on the mean of the chosen probabilities (cpr_head.py:833-837)
    nr_must = rows_equal & (sm > bound)
    nbad = int((nr.bool() != ora['not_refine'].bool())[nr_must].sum())
    assert nbad == 0, f'{what}: {nbad} not_refine flips with margin > {bound:.1e}'
    rows = rows_equal & (nr.bool() == ora['not_refine'].bool())
    assert_close(pts[rows], ora['refine_pts'][rows], 1e-4, what + ' refined points')
    assert_close(sc[rows], ora['refine_scores'][rows], 1e-4, what + ' refine scores')
    stats = dict(samples=int(G * K), within_bound=int((~decisive).sum()), flips_within_bound=int(flips.sum()),
                 gts=int(G), gts_with_flips=int((~rows_equal).sum()), not_refine_within_bound=int((rows_equal & ~(sm > bound)).sum()),
                 not_refine_flips=int((nr.bool() != ora['not_refine'].bool()).sum()), bound=float(bound))
    print(f'[margin harness] {what}: {stats}')
    return stats


def nms_replay_per_class(boxes, scores, labels, iou_thr, max_num):
    """exact greedy NMS replayed class by class (numpy): valid whenever the class-offset trick of mmcv's batched_nms keeps the classes
    disjoint (true for non-negative-ish coordinates).  returns keep indices (into the candidate list) in descending
    score order, cut to max_num — the same contract as oracle.p2p.multiclass_nms's `keep`."""
    from oracle import p2p as op2p
    keep_all = []
    for c in np.unique(labels):
        idx = np.nonzero(labels == c)[0]
        k = op2p.nms(torch.from_numpy(boxes[idx]), torch.from_numpy(scores[idx]), iou_thr).numpy()
        keep_all.append(idx[k])
    keep = np.concatenate(keep_all) if keep_all else np.zeros(0, np.int64)
    order = np.argsort(-scores[keep], kind='stable')
    keep = keep[order]
    return keep[:max_num] if max_num > 0 else keep
