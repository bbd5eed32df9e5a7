"""ptb_rpn_proposals on 16 tiles x 81 840 anchors (BASELINE.json configs[3]): CUDA-event time of the whole call with the round-2 bitmask NMS
and with the round-1 serial NMS (PTB_RPN_NMS=serial); argv[1] == 'ncu': three calls for a launch list."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import anchors as oa
from pointtinybenchmark_b200 import ops
from pointtinybenchmark_b200.rpn import AnchorGenerator
dev = torch.device('cuda:0')
c = oa.RPN_CFG
cls, box, shapes = oa.synth_rpn_inputs(3, B=16, size=(512, 640), strides=c['strides'])
ag = AnchorGenerator(scales=c['scales'], ratios=c['ratios'], strides=c['strides'])
img_hw = torch.tensor([[s[0], s[1]] for s in shapes], dtype=torch.int32, device=dev)
cl = [t.to(dev) for t in cls]; bx = [t.to(dev) for t in box]; ba = torch.stack(ag.base_anchors).to(dev)


def f():
    return ops.rpn_proposals(cl, bx, ba, ag.strides, img_hw, c['means'], c['stds'], 16 / 1000, 1000, c['min_bbox_size'], c['iou_threshold'], 1000)


if len(sys.argv) > 1 and sys.argv[1] == 'ncu':
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    sys.exit(0)
res = {'tiles': len(shapes)}
for mode in ('bitmask', 'serial'):
    if mode == 'serial':
        os.environ['PTB_RPN_NMS'] = 'serial'
    else:
        os.environ.pop('PTB_RPN_NMS', None)
    for _ in range(3):
        f()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(20):
        f()
    b.record(); torch.cuda.synchronize()
    res[mode + '_ms_per_16_tiles'] = a.elapsed_time(b) / 20
print(json.dumps(res))
