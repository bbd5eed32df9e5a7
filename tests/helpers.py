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
