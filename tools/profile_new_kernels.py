"""Workload for `ncu --set full` captures of the kernels added in session 2 of round 1 (Hungarian matching, RPN proposals,
MaxIoUAssigner):   ncu --set full --clock-control none --import-source on -k regex:'hungarian_v2_kernel|rpn_|miou_|lsap_prep' \
                       -o gpurun_out/new_kernels python tools/profile_new_kernels.py
One warm-up pass (not profiled when -s is used) and one measured pass of each op at its BASELINE.json shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                              # noqa: E402  (synthetic input generators only)
from pointtinybenchmark_b200 import ops   # noqa: E402
from pointtinybenchmark_b200.rpn import AnchorGenerator   # noqa: E402

dev = torch.device('cuda:0')
passes = int(sys.argv[1]) if len(sys.argv) > 1 else 1
g = torch.Generator().manual_seed(9)
Bh, H, W, N, nh = 16, 100, 168, 80, 100
Qh = H * W
clsh = (torch.randn(Bh, Qh, N, generator=g) * 1.5 - 3.0).to(dev)
xs, ys = (torch.arange(Qh) % W).float() * 8, (torch.arange(Qh) // W).float() * 8
prop = (torch.stack([xs, ys], 1)[None] + torch.randn(Bh, Qh, 2, generator=g) * 4).to(dev).contiguous()
gts = (torch.rand(Bh, nh, 2, generator=g) * torch.tensor([1333., 800.])).to(dev)
gl = torch.randint(0, N, (Bh, nh), generator=g).int().to(dev)
cost = torch.empty(Bh * Qh * nh, device=dev)
gi = torch.zeros(Bh * Qh, dtype=torch.int64, device=dev)
cls4, box4, shp4 = bench.synth_rpn_outputs(21, 16)
c4 = dict(scales=[2], ratios=[0.5, 1.0, 2.0], strides=[4, 8, 16, 32, 64], means=(0., 0., 0., 0.), stds=(1., 1., 1., 1.))
ag = AnchorGenerator(scales=c4['scales'], ratios=c4['ratios'], strides=c4['strides'])
cls4, box4 = [t.to(dev) for t in cls4], [t.to(dev) for t in box4]
base = torch.stack(ag.base_anchors).to(dev)
ihw = torch.tensor([[s[0], s[1]] for s in shp4], dtype=torch.int32, device=dev)
a4, g4, l4, i4 = [t.to(dev) for t in bench.synth_dense_anchors(11)]
for _ in range(passes):
    for b in range(Bh):
        ops.p2p_cost_matrix(clsh[b], prop[b], None, gts[b], gl[b], 2.0, 0.25, 2.0, 1e-12, 0.1, 1333.0, 800.0, out=cost[b * Qh * nh:(b + 1) * Qh * nh])
    gi.zero_()
    st = ops.hungarian_v2_batch(cost, [(Qh, nh)] * Bh, 5, gi, [b * Qh for b in range(Bh)])
    ops.rpn_proposals(cls4, box4, base, ag.strides, ihw, c4['means'], c4['stds'], 16 / 1000, 1000, 0, 0.7, 1000)
    ops.max_iou_assign(a4, g4, l4, i4, pos_iou_thr=0.7, neg_iou_thr=0.3, min_pos_iou=0.3, match_low_quality=True, ignore_iof_thr=0.5)
    torch.cuda.synchronize()
print('status', st.cpu().tolist(), 'assigned', int((gi > 0).sum()))
