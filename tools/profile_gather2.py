"""neighbor gather (C = 256 reference data flow, C = 160 training logits, C = 80) at the headline shape: TMA-staged window vs the LDG kernel
(PTB_GATHER_TMA=0), CUDA events, L2 flushed between launches; bit-equality of the two; `ncu` mode = three C=256 launches."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointtinybenchmark_b200 import ops

dev = torch.device('cuda:0')
torch.manual_seed(0)
B, H, W, n, r, s = 8, 100, 168, 500, 8, 8
centers = (torch.rand(B * n, 2, device=dev) * torch.tensor([1344., 800.], device=dev)).contiguous()
bag_img = torch.arange(B, device=dev, dtype=torch.int32).repeat_interleave(n).contiguous()
pad_hw = torch.tensor([[800, 1344]] * B, dtype=torch.int32, device=dev)
off = ops.circle_offsets(r, s).to(dev)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
maps = {C: torch.randn(B, H, W, C, device=dev) for C in (256, 160, 80)}


def ktime(fn, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.add_(1.0)
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sum(ts) / len(ts)


if len(sys.argv) > 1 and sys.argv[1] == 'ncu':
    for _ in range(3):
        ops.bag_gather(maps[256], centers, bag_img, off, s, pad_hw)
    torch.cuda.synchronize()
    sys.exit(0)
out = {}
for C, m in maps.items():
    res = {}
    for mode in ('1', '0'):
        os.environ['PTB_GATHER_TMA'] = mode
        t = ktime(lambda: ops.bag_gather(m, centers, bag_img, off, s, pad_hw))
        res[mode] = ops.bag_gather(m, centers, bag_img, off, s, pad_hw)
        alg = m.numel() * 4 + B * n * off.shape[0] * (8 + C * 4 + 1)
        out[f'C{C}_{"tma" if mode == "1" else "ldg"}'] = dict(ms=t, gb_per_s=alg / t / 1e6)
    out[f'C{C}_bit_identical'] = all(torch.equal(a, b) for a, b in zip(res['1'], res['0']))
os.environ['PTB_GATHER_TMA'] = '1'
os.environ['PTB_GATHER_CC'] = '32'
out['C256_tma_cc32'] = dict(ms=ktime(lambda: ops.bag_gather(maps[256], centers, bag_img, off, s, pad_hw)))
os.environ.pop('PTB_GATHER_CC', None)
os.environ.pop('PTB_GATHER_TMA', None)
print(json.dumps(out))
