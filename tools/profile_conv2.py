"""conv3x3 256->256 fp16-split kernel at the headline shape in the three cluster modes (PTB_CONV_CLUSTER 1 / 2 / 3): CUDA-event timing with
L2 flushed between launches, or (argv[1] == 'ncu') three launches of the mode in PTB_CONV_CLUSTER for `ncu --set full -k regex:conv_tc`."""
import os, sys, json
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
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def run():
    return ops.conv3x3_c256_f16(h16, l16, wh16, wl16, invw, dinv)


if len(sys.argv) > 1 and sys.argv[1] == 'ncu':
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    sys.exit(0)
out, ys = {}, {}
for mode in ('1', '2', '3', '3_nosplit'):
    os.environ['PTB_CONV_CLUSTER'] = mode[0]
    os.environ['PTB_CONV_TAIL_SPLIT'] = '0' if mode.endswith('nosplit') else '1'
    for _ in range(3):
        run()
    ts = []
    for _ in range(20):
        flush.add_(1.0)
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); y, st = run(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    # back-to-back (no flush): what the tower sees
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(20):
        run()
    b.record(); torch.cuda.synchronize()
    out[mode] = dict(mean_ms=sum(ts) / len(ts), min_ms=min(ts), back_to_back_ms=a.elapsed_time(b) / 20)
    ys[mode] = (y.clone(), st.clone())
out['bit_identical_3_vs_3_nosplit'] = bool(torch.equal(ys['3'][0], ys['3_nosplit'][0]))
out['bit_identical_1_vs_3'] = bool(torch.equal(ys['1'][0], ys['3'][0]))
out['stats_rel_diff_1_vs_3'] = float(((ys['1'][1] - ys['3'][1]).abs() / ys['1'][1].abs().clamp(min=1e-30)).max())
print(json.dumps(out))
