"""two CPRHead training steps (forward towers + loss + backward) at the headline shape: target for the ncu launch list
(`ncu --metrics gpu__time_duration.sum`) and for `ncu --set full` of the loss-path kernels."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pointtinybenchmark_b200 import cpr_head  # noqa: F401
from pointtinybenchmark_b200.registry import build_head

dev = torch.device('cuda:0')
head = build_head(bench.head_cfg()).to(dev).train()
sd = head.state_dict(); sd.update(bench.head_weights()); head.load_state_dict(sd)
x, gtb, gtl, aid, metas = bench.synth_batch(bench.CFG['B'], 1234)
xg = x.to(dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
gtb = [t.to(dev) for t in gtb]; gtl = [t.to(dev) for t in gtl]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for i in range(n):
    head.zero_grad(set_to_none=True)
    cf, inf = head((xg,))
    losses = head.loss(cf, inf, gtb, gtl, metas)
    sum(v for k, v in losses.items() if 'loss' in k).backward()
torch.cuda.synchronize()
print({k: float(v.reshape(-1)[0]) for k, v in losses.items()})
