"""Per-kernel CUDA-event timings of the CPR point path at the headline config (scratch tool; bench.py is the contract)."""
import json
import sys
import os
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointtinybenchmark_b200 import ops  # noqa: E402


def timeit(fn, n=20, warm=3, flush=None):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        if flush is not None:
            flush.add_(1.0)
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    B, H, W, C, n, r, s, ncls = 8, 100, 168, 256, 500, 8, 8, 80
    feat = torch.relu(torch.randn(B, H, W, C, device=dev))
    centers = (torch.rand(B * n, 2, device=dev) * torch.tensor([1344., 800.], device=dev)).contiguous()
    bag_img = torch.arange(B, device=dev, dtype=torch.int32).repeat_interleave(n).contiguous()
    labels = torch.randint(0, ncls, (B * n,), device=dev, dtype=torch.int32)
    img_ptr = (torch.arange(B + 1, device=dev) * n).int()
    pad_hw = torch.tensor([[800, 1344]] * B, dtype=torch.int32, device=dev)
    img_hw = torch.tensor([[800, 1333]] * B, dtype=torch.int32, device=dev)
    off = ops.circle_offsets(r, s).to(dev)
    K = off.shape[0]
    wc = torch.randn(ncls, C, device=dev) * 0.05
    bc = torch.zeros(ncls, device=dev)
    w2 = torch.randn(2 * ncls, C, device=dev) * 0.05
    b2 = torch.zeros(2 * ncls, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)   # 256 MB > L2
    res = {}
    G = B * n
    out_bytes = G * K * C * 4
    alg = B * H * W * C * 4 + G * K * 8 + out_bytes + G * K
    med, best = timeit(lambda: ops.bag_gather(feat, centers, bag_img, off, s, pad_hw), flush=flush)
    res['gather256'] = dict(ms=med, best_ms=best, alg_GBs=alg / med / 1e6, alg_bytes=alg)
    lmap = ops.linear_rows(feat.reshape(-1, C), wc, bc).reshape(B, H, W, ncls)
    med, best = timeit(lambda: ops.linear_rows(feat.reshape(-1, C), wc, bc), flush=flush)
    res['linear80'] = dict(ms=med, best_ms=best, tflops=2 * B * H * W * C * ncls / med / 1e9)
    med, best = timeit(lambda: ops.linear_rows(feat.reshape(-1, C), w2, b2), flush=flush)
    res['linear160'] = dict(ms=med, best_ms=best, tflops=2 * B * H * W * C * 2 * ncls / med / 1e9)
    med, best = timeit(lambda: ops.bag_gather(lmap, centers, bag_img, off, s, pad_hw, pts=False, valid=False), flush=flush)
    res['gather80'] = dict(ms=med, best_ms=best, alg_GBs=(B * H * W * ncls * 4 + G * K * ncls * 4) / med / 1e6)
    med, best = timeit(lambda: ops.neg_mask(B, H, W, s, pad_hw, centers, labels, img_ptr, s * r, ncls, True), flush=flush)
    res['neg_mask'] = dict(ms=med, best_ms=best)
    groups = ops.label_groups(bag_img, labels, ncls)
    rc = ops._refine_cfg(0.1, 0.5, 0.1, True, True, False)
    med, best = timeit(lambda: ops.refine_fused(lmap, ncls, centers, labels, bag_img, off, s, pad_hw, img_hw, groups, rc), flush=flush)
    res['refine_fused'] = dict(ms=med, best_ms=best)
    med, best = timeit(lambda: ops.label_groups(bag_img, labels, ncls), flush=flush)
    res['label_groups(torch)'] = dict(ms=med, best_ms=best)
    x = torch.randn(B, C, H, W, device=dev).contiguous(memory_format=torch.channels_last)
    conv = torch.nn.Conv2d(C, C, 3, padding=1, bias=False).to(dev).to(memory_format=torch.channels_last)
    gn = torch.nn.GroupNorm(32, C).to(dev)
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        with torch.no_grad():
            med, best = timeit(lambda: torch.relu(gn(conv(x))), n=10, flush=flush)
        res[f'torch_conv_gn_relu_tf32={tf32}'] = dict(ms=med, best_ms=best, tflops=2 * 9 * C * C * B * H * W / med / 1e9)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
