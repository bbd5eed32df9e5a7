"""CPR output -> annotation file: the on-disk hand-off that turns refined points into the next stage's training set
(reference docs/cpr/README.md:81-103).

    head.get_bboxes rows [x1, y1, x2, y2, score, ann_id(, geo...)]
      -> bbox2result   (mmdet/core/bbox/transforms.py:124-140: rows split per class, numpy)
      -> det2json      (mmdet/datasets/coco.py:212-234 CocoDataset._det2json: xywh boxes, score, category id, ann_id = int(col 5),
                        geo = columns 6.. that are >= 0 rounded to 0.1)
      -> result2ann    (exp/tools/result2ann.py:55-93: refined boxes written back into the original annotation file)

Host-side format code only (no kernels).  bbox2result / det2json are pinned to outputs of the real reference functions
(tests/golden/cpr_result_json.json, written by oracle/make_golden.py).  result2ann depends on pycocotools' COCO.loadRes (absent in
this image and not part of /root/reference): its published behaviour for bbox results (area = w*h, a 4-corner polygon as
segmentation, id = position + 1, iscrowd = 0) is restated here and is NOT pinned by a reference run — parity unpinned for that
function.

Quirk preserved: ann_id travels in a float32 column (cpr_head.py:1269), so ids above 2**24 are rounded before int().
"""
import copy
import json

import numpy as np
import torch


def bbox2result(bboxes, labels, num_classes):
    """split detection rows per class (mmdet/core/bbox/transforms.py:124-140): list of num_classes numpy arrays.
    An image without detections yields (0, 5) arrays whatever the row width — the reference's behaviour."""
    if len(bboxes) == 0:
        return [np.zeros((0, 5), dtype=np.float32) for _ in range(num_classes)]
    rows = bboxes.detach().cpu().numpy() if isinstance(bboxes, torch.Tensor) else np.asarray(bboxes)
    cls = labels.detach().cpu().numpy() if isinstance(labels, torch.Tensor) else np.asarray(labels)
    return [rows[cls == c] for c in range(num_classes)]


def xyxy2xywh(bbox):
    """corner box -> COCO [x, y, w, h] as python floats (mmdet/datasets/coco.py:176-194)."""
    x1, y1, x2, y2 = (v for v in np.asarray(bbox).tolist()[:4])
    return [x1, y1, x2 - x1, y2 - y1]


def det2json(results, img_ids, cat_ids):
    """per-image, per-class detection arrays -> COCO result records, field for field what CocoDataset._det2json emits
    (mmdet/datasets/coco.py:212-234): `bbox` xywh, `score` = column 4, `ann_id` = int(column 5) when present, `geo` = the
    non-negative entries of columns 6.. rounded to one decimal (the -1 padding of fill_list_to_tensor drops out)."""
    records = []
    for img_id, per_class in zip(img_ids, results):
        for cat_id, dets in zip(cat_ids, per_class):
            for row in dets:
                rec = {'image_id': img_id, 'bbox': xyxy2xywh(row), 'score': float(row[4]), 'category_id': cat_id}
                if len(row) >= 6:
                    rec['ann_id'] = int(row[5])
                if len(row) >= 7:
                    geo = [round(v, 1) for v in row[6:] if v >= 0]
                    if len(geo) % 2:
                        raise AssertionError('geo must hold (x, y) pairs')
                    rec['geo'] = geo
                records.append(rec)
    return records


def _json_default(o):
    if isinstance(o, (np.floating,)):
        return float(o)
    if isinstance(o, (np.integer,)):
        return int(o)
    raise TypeError(type(o))


def head_results_to_json(det_results, num_classes, img_ids, cat_ids, out_file=None):
    """[(det (n, >=6), labels (n,))] per image (CPRHead.get_bboxes / simple_test) -> COCO-style result list (+ file)."""
    per_class = [bbox2result(d, l, num_classes) for d, l in det_results]
    js = det2json(per_class, img_ids, cat_ids)
    if out_file is not None:
        with open(out_file, 'w') as f:
            json.dump(js, f, default=_json_default)
    return js


def turn_bbox_wh(bbox, new_wh):
    """Annotation box [x, y, w, h] re-sized around its own centre (what exp/tools/result2ann.py:43-53 does to a refined pseudo box when the
    caller asks for a fixed size); a non-positive size leaves the box as it is.  The arithmetic is the tool's: centre = corner + size / 2,
    new corner = centre - new size / 2 (Python floats, so the json it writes has the same digits)."""
    new_w, new_h = new_wh
    if not (new_w > 0 and new_h > 0):
        return bbox
    x, y, w, h = bbox
    centre = (x + w / 2, y + h / 2)
    out = [centre[0] - new_w / 2, centre[1] - new_h / 2, new_w, new_h]
    moved = (out[0] + out[2] / 2, out[1] + out[3] / 2)
    if (round(moved[0]), round(moved[1])) != (round(centre[0]), round(centre[1])):     # float cancellation would have to be enormous
        raise AssertionError(f'resizing {bbox} to {new_wh} moved its centre from {centre} to {moved}')
    return out


def result2ann(ori_dataset, det_json, wh=-1):
    """exp/tools/result2ann.py:55-93 without pycocotools: returns a copy of the COCO dataset dict whose annotations carry the
    refined pseudo boxes (and geo) of det_json, matched by ann_id.  (COCO.loadRes restated, see module docstring.)"""
    ds = copy.deepcopy(ori_dataset)
    by_id = {a['id']: a for a in ds['annotations']}
    if isinstance(wh, (int, float)):
        wh = (wh, wh)
    for res in det_json:
        ori = by_id[res['ann_id']]
        assert ori['id'] == res['ann_id'], f'{ori} vs {res}'
        for key in ('image_id', 'category_id'):
            assert ori[key] == res[key], key
        assert ori.get('iscrowd', 0) == 0, 'iscrowd'           # loadRes stamps iscrowd = 0 on every result
        bb = res['bbox']
        x1, x2, y1, y2 = bb[0], bb[0] + bb[2], bb[1], bb[1] + bb[3]
        ori['bbox'] = turn_bbox_wh(bb, wh)
        ori['segmentation'] = res.get('segmentation', [[x1, y1, x1, y2, x2, y2, x2, y1]])
        ori['area'] = bb[2] * bb[3]
        if 'geo' in res:
            ori['geo'] = res['geo']
    return ds
