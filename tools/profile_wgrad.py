"""wgrad (dW of a conv3x3 256->256) at the headline shape: single-CTA vs CTA-pair kernel (PTB_WGRAD_PAIR=0 / 1), CUDA-event timing of the whole
op (tensor-core partials + fixed-order reduce) with L2 flushed between launches, bit-equality of the two modes; argv[1] == 'ncu': three
launches for `ncu --set full -k regex:wgrad_tc`."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointtinybenchmark_b200 import ops
dev = torch.device('cuda:0')
torch.manual_seed(0)
B, H, W, C = 8, 100, 168, 256
x = torch.relu(torch.randn(B, H, W, C, device=dev))
dy = torch.randn(B, H, W, C, device=dev) * 1e-3
xh, xl, xinv = ops.split_f16(x, auto_scale=True)
dh, dl, dinv = ops.split_f16(dy, auto_scale=True)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def run():
    return ops.conv_tc_wgrad_f16(dh, dl, xh, xl, 9, 1.0, dinv, xinv)


if len(sys.argv) > 1 and sys.argv[1] == 'ncu':
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    sys.exit(0)
out, ys = {}, {}
for mode in ('0', '1'):
    os.environ['PTB_WGRAD_PAIR'] = mode
    for _ in range(3):
        run()
    ts = []
    for _ in range(20):
        flush.add_(1.0)
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); y = run(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    out['pair' if mode == '1' else 'single'] = dict(mean_ms=sum(ts) / len(ts), min_ms=min(ts))
    ys[mode] = y.clone()
out['bit_identical'] = bool(torch.equal(ys['0'], ys['1']))
ref = torch.nn.grad.conv2d_weight(x.permute(0, 3, 1, 2).double(), (256, 256, 3, 3), dy.permute(0, 3, 1, 2).double(), padding=1)
out['max_rel_err_vs_fp64'] = float((ys['1'].double() - ref).abs().max() / ref.abs().max())
print(json.dumps(out))
