"""ptb_p2p_cost_matrix + ptb_hungarian_v2_batch (topk_k 5) at 16 x (16 800 proposals x 100 GTs) — the bench's `p2p_hungarian` case on its
own: CUDA-event time of the assignment, and of the Hungarian launch alone; argv[1] == 'ncu': two calls for a launch list."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointtinybenchmark_b200 import ops
dev = torch.device('cuda:0')
g3 = torch.Generator().manual_seed(9)
H, W, N, stride = 100, 168, 80, 8
Bh, Qh, nh = 16, H * W, 100
clsh = (torch.randn(Bh, Qh, N, generator=g3) * 1.5 - 3.0).to(dev)
xs = (torch.arange(Qh) % W).float() * stride
ys = (torch.arange(Qh) // W).float() * stride
prop = (torch.stack([xs, ys], 1)[None] + torch.randn(Bh, Qh, 2, generator=g3) * 4).to(dev).contiguous()
gts_h = (torch.rand(Bh, nh, 2, generator=g3) * torch.tensor([1333., 800.])).to(dev)
gl_h = torch.randint(0, N, (Bh, nh), generator=g3).int().to(dev)
cost_flat = torch.empty(Bh * Qh * nh, device=dev)
gi_out = torch.zeros(Bh * Qh, dtype=torch.int64, device=dev)
shapes_h = [(Qh, nh)] * Bh
for b in range(Bh):
    ops.p2p_cost_matrix(clsh[b], prop[b], None, gts_h[b], gl_h[b], 2.0, 0.25, 2.0, 1e-12, 0.1, 1333.0, 800.0, out=cost_flat[b * Qh * nh:(b + 1) * Qh * nh])


def solve():
    gi_out.zero_()
    return ops.hungarian_v2_batch(cost_flat, shapes_h, 5, gi_out, [b * Qh for b in range(Bh)])


if len(sys.argv) > 1 and sys.argv[1] == 'ncu':
    solve(); solve()
    torch.cuda.synchronize()
    sys.exit(0)
for _ in range(2):
    st = solve()
a, b = torch.cuda.Event(True), torch.cuda.Event(True)
a.record()
for _ in range(5):
    solve()
b.record(); torch.cuda.synchronize()
print(json.dumps(dict(hungarian_ms_per_batch16=a.elapsed_time(b) / 5, status_ok=bool(int(st.max()) == 0), matched=int((gi_out > 0).sum()))))
