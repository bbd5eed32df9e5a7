"""runs the neighbor-gather (C=256) and the fused refine a few times at the headline config: target for `ncu --set full`."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointtinybenchmark_b200 import ops

dev = torch.device('cuda:0')
torch.manual_seed(0)
B, H, W, C, n, r, s, ncls = 8, 100, 168, 256, 500, 8, 8, 80
feat = torch.relu(torch.randn(B, H, W, C, device=dev))
centers = (torch.rand(B * n, 2, device=dev) * torch.tensor([1344., 800.], device=dev)).contiguous()
bag_img = torch.arange(B, device=dev, dtype=torch.int32).repeat_interleave(n).contiguous()
labels = torch.randint(0, ncls, (B * n,), device=dev, dtype=torch.int32)
pad_hw = torch.tensor([[800, 1344]] * B, dtype=torch.int32, device=dev)
img_hw = torch.tensor([[800, 1333]] * B, dtype=torch.int32, device=dev)
off = ops.circle_offsets(r, s).to(dev)
lmap = torch.randn(B, H, W, ncls, device=dev)
groups = ops.label_groups(bag_img, labels, ncls)
rc = ops._refine_cfg(0.1, 0.5, 0.1, True, True, False)
wc = torch.randn(ncls, C, device=dev) * 0.05
h16, l16, dinv = ops.split_f16(feat, auto_scale=True)
convw = torch.randn(256, C, 3, 3, device=dev) * 0.02
wh, wl, invw = ops.conv3x3_pack_weight_f16(convw)
bc = torch.zeros(ncls, device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    ops.bag_gather(feat, centers, bag_img, off, s, pad_hw)
    ops.refine_fused(lmap, ncls, centers, labels, bag_img, off, s, pad_hw, img_hw, groups, rc)
    ops.linear_rows(feat.reshape(-1, C), wc, bc)
    ops.conv3x3_c256_f16(h16, l16, wh, wl, invw, dinv)
torch.cuda.synchronize()
