"""Multi-GPU plumbing of the head: images are the independent units, every rank runs the whole head on its own
images, inference needs no collective at all.  Training has exactly one exchange step — the gradient all-reduce the reference gets
from MMDistributedDataParallel (mmdet/apis/train.py:75-86) — done here for the head's parameters as ONE flat fp32 bucket over
NCCL (9.6 MB for the shipped head: a single NVLink/NVSwitch all-reduce, latency-bound, so no bucketing by layer).  The loss
normalisers stay rank-local like the reference's (cpr_head.py:1180, 1227; no reduce_mean)."""
import torch
import torch.distributed as dist


def allreduce_grads(module, group=None, average=True):
    """all-reduce (sum, then / world) the gradients of `module`'s parameters in place through one flat bucket.
    Parameters without a gradient on this rank contribute zeros (DDP's find_unused_parameters semantics).  Returns the number
    of bytes exchanged per rank (0 when not distributed)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    params = [p for p in module.parameters() if p.requires_grad]
    if not params:
        return 0
    dev, dt = params[0].device, torch.float32
    sizes = [p.numel() for p in params]
    flat = torch.zeros(sum(sizes), dtype=dt, device=dev)
    o = 0
    for p, n in zip(params, sizes):
        if p.grad is not None:
            flat[o:o + n].copy_(p.grad.reshape(-1))
        o += n
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat.div_(dist.get_world_size(group))
    o = 0
    for p, n in zip(params, sizes):
        g = flat[o:o + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        o += n
    return flat.numel() * flat.element_size()


def parse_losses(losses, group=None):
    """BaseDetector._parse_losses (mmdet/models/detectors/base.py:179-212): mean every entry (lists are summed), total `loss` = the sum
    of the entries whose key contains 'loss', and the logged values averaged over the ranks.  The reference issues one all-reduce and
    one `.item()` host sync PER scalar (4-6 per iteration); here the scalars travel as ONE packed tensor and come back with one copy.
    returns (loss tensor for backward — rank-local, like the reference —, {name: float})."""
    from collections import OrderedDict
    log_vars = OrderedDict()
    for k, v in losses.items():
        if isinstance(v, torch.Tensor):
            log_vars[k] = v.mean()
        elif isinstance(v, list):
            log_vars[k] = sum(x.mean() for x in v)
        else:
            raise TypeError(f'{k} is not a tensor or list of tensors')
    loss = sum(v for k, v in log_vars.items() if 'loss' in k)
    log_vars['loss'] = loss
    packed = torch.stack([v.detach().float().reshape(()) for v in log_vars.values()])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        packed = packed / dist.get_world_size(group)
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    vals = packed.cpu().tolist()
    return loss, OrderedDict(zip(log_vars.keys(), vals))
