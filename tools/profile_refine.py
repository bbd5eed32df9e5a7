"""times ptb_cpr_refine_fused at the headline shape (8 x 100x168x80 logit map, 4000 GTs, K = 289) with the TMA-staged window and with
the global-memory path (PTB_REFINE_TMA=0), CUDA events, L2 flushed between launches; also the ncu target for the kernel."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointtinybenchmark_b200 import ops

dev = torch.device('cuda:0')
torch.manual_seed(0)
B, H, W, n, r, s, ncls = 8, 100, 168, 500, 8, 8, 80
centers = (torch.rand(B * n, 2, device=dev) * torch.tensor([1344., 800.], device=dev)).contiguous()
bag_img = torch.arange(B, device=dev, dtype=torch.int32).repeat_interleave(n).contiguous()
labels = torch.randint(0, ncls, (B * n,), device=dev, dtype=torch.int32)
pad_hw = torch.tensor([[800, 1344]] * B, dtype=torch.int32, device=dev)
img_hw = torch.tensor([[800, 1333]] * B, dtype=torch.int32, device=dev)
off = ops.circle_offsets(r, s).to(dev)
lmap = torch.randn(B, H, W, ncls, device=dev) * 1.5 - 2.0
groups = ops.label_groups(bag_img, labels, ncls)
rc = ops._refine_cfg(0.1, 0.5, 0.1, True, True, False)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def run():
    return ops.refine_fused(lmap, ncls, centers, labels, bag_img, off, s, pad_hw, img_hw, groups, rc, want_chosen=True)


def ktime(n=20):
    for _ in range(3):
        run()
    ts = []
    for _ in range(n):
        flush.add_(1.0)
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); run(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sum(ts) / len(ts), min(ts)


if len(sys.argv) > 1 and sys.argv[1] == 'ncu':
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    sys.exit(0)
out = {}
res = {}
for mode in ('1', '0'):
    os.environ['PTB_REFINE_TMA'] = mode
    out['tma' if mode == '1' else 'global'] = dict(zip(('mean_ms', 'min_ms'), ktime()))
    res[mode] = run()
same = all(torch.equal(a, b) for a, b in zip(res['1'], res['0']))
out['staged_equals_global_path_bitwise'] = same
print(json.dumps(out))
