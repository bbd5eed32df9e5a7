"""CPR output -> annotation file: bbox2result / det2json / result2ann of pointtinybenchmark_b200.results.
The json fixture was produced by the REAL reference (head.get_bboxes with out_geo -> mmdet bbox2result -> CocoDataset._det2json,
oracle/make_golden.py::golden_result_json)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import cpr as ocpr, synth
from pointtinybenchmark_b200 import results as R
from tests.helpers import oracle_cfg


def _oracle_dets(gold):
    inp = synth.cpr_inputs('lite', 1234)
    cfg = oracle_cfg(inp['cfgd'])
    metas = [dict(m, scale_factor=gold['scale_factor']) for m in inp['img_metas']]
    res = ocpr.cpr_get_bboxes(inp['cls_feat'], inp['weights'], inp['gt_bboxes'], inp['gt_labels'], inp['gt_anns_id'], metas, cfg,
                              rescale=True, out_geo=True)
    return inp, metas, res


def test_det2json_matches_the_reference(golden_dir, tmp_path):
    gold = json.load(open(os.path.join(golden_dir, 'cpr_result_json.json')))
    inp, _, res = _oracle_dets(gold)
    out = tmp_path / 'res.json'
    js = R.head_results_to_json(res, inp['cfgd']['num_classes'], gold['img_ids'], gold['cat_ids'], out_file=str(out))
    assert json.loads(out.read_text()) == gold['results']
    assert len(js) == sum(len(l) for l in inp['gt_labels'])
    r0 = js[0]
    assert set(r0) == {'image_id', 'bbox', 'score', 'category_id', 'ann_id', 'geo'} and isinstance(r0['ann_id'], int)
    assert len(r0['geo']) % 2 == 0 and all(g >= 0 for g in r0['geo'])


def test_bbox2result_shapes_and_empty():
    det = torch.tensor([[0., 0, 1, 1, .5, 7], [2, 2, 3, 3, .25, 8], [4, 4, 5, 5, .75, 9]])
    lab = torch.tensor([2, 0, 2])
    out = R.bbox2result(det, lab, 4)
    assert [len(o) for o in out] == [1, 0, 2, 0] and out[2][1][5] == 9
    empty = R.bbox2result(torch.zeros(0, 6), torch.zeros(0, dtype=torch.long), 3)
    assert [o.shape for o in empty] == [(0, 5)] * 3                       # the reference's fixed (0, 5) quirk
    js = R.det2json([out], [42], [1, 2, 3, 4])
    assert [r['category_id'] for r in js] == [1, 3, 3] and js[0]['bbox'] == [2.0, 2.0, 1.0, 1.0]
    big = R.det2json([[np.array([[0, 0, 1, 1, .5, 2 ** 24 + 1]], dtype=np.float32)]], [1], [1])
    assert big[0]['ann_id'] == 2 ** 24                                    # float32 column rounds the id (cpr_head.py:1269)


def test_result2ann_writes_refined_boxes_back():
    ds = dict(images=[dict(id=5)], categories=[dict(id=3)],
              annotations=[dict(id=11, image_id=5, category_id=3, iscrowd=0, bbox=[10, 10, 16, 16], area=256, segmentation=[]),
                           dict(id=12, image_id=5, category_id=3, iscrowd=0, bbox=[40, 40, 16, 16], area=256, segmentation=[])])
    det = [dict(image_id=5, category_id=3, ann_id=12, score=0.9, bbox=[44.0, 38.0, 16.0, 16.0], geo=[52.0, 46.0, 50.0, 44.0])]
    out = R.result2ann(ds, det, wh=-1)
    a12 = [a for a in out['annotations'] if a['id'] == 12][0]
    assert a12['bbox'] == [44.0, 38.0, 16.0, 16.0] and a12['area'] == 256.0 and a12['geo'] == det[0]['geo']
    assert a12['segmentation'] == [[44.0, 38.0, 44.0, 54.0, 60.0, 54.0, 60.0, 38.0]]
    assert [a for a in out['annotations'] if a['id'] == 11][0]['bbox'] == [10, 10, 16, 16]     # untouched
    assert ds['annotations'][1]['bbox'] == [40, 40, 16, 16]                                    # input not mutated
    out32 = R.result2ann(ds, det, wh=32)
    assert [a for a in out32['annotations'] if a['id'] == 12][0]['bbox'] == [36.0, 30.0, 32, 32]
    with pytest.raises(AssertionError):
        R.result2ann(ds, [dict(det[0], category_id=4)])
