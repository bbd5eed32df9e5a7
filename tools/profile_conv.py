"""runs the fp16x2 and 3xTF32 conv kernels at the headline shape: target for `ncu --set full -k regex:conv3x3`."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointtinybenchmark_b200 import ops
dev = torch.device('cuda:0')
torch.manual_seed(0)
B, H, W, C = 8, 100, 168, 256
x = torch.relu(torch.randn(B, H, W, C, device=dev))
w = torch.randn(256, C, 3, 3, device=dev) * 0.02
h16, l16, dinv = ops.split_f16(x, auto_scale=True)
wh16, wl16, invw = ops.conv3x3_pack_weight_f16(w)
xh, xl = ops.split_tf32(x)
wh, wl = ops.conv3x3_pack_weight(w)
for _ in range(3):
    ops.conv3x3_c256_f16(h16, l16, wh16, wl16, invw, dinv)
    ops.conv3x3_c256(xh, xl, wh, wl)
torch.cuda.synchronize()
