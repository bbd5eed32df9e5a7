"""Host-side layer containers with the reference's parameter names (checkpoint compatibility).

Conv towers: at inference they run on the hand-written tcgen05 implicit-GEMM kernel of csrc/conv_tc.cu
(fp16 two-term split by default, 3xTF32 with PTB_CONV_MODE=tf32x3; both fp32-accurate); under autograd (training) the shipped
256 -> 256 geometry runs `_TowerTCFn`: the same forward kernel plus hand-written GroupNorm/ReLU backward, dgrad (the forward
kernel on transposed, flipped weights) and a tcgen05 wgrad with MN-major operands.  Other geometries (or PTB_TOWER_TRAIN=cudnn)
use cuDNN fp32 through torch with TF32 switched off locally (library path).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class ConvModule(nn.Module):
    """conv3x3 (bias iff no norm) -> GN/BN -> ReLU with mmcv.cnn.ConvModule's submodule names (`conv`, `gn`|`bn`)."""

    def __init__(self, cin, cout, k=3, stride=1, padding=1, norm_cfg=None, bias='auto', act=True):
        super().__init__()
        with_norm = norm_cfg is not None
        if bias == 'auto':
            bias = not with_norm
        self.conv = nn.Conv2d(cin, cout, k, stride, padding, bias=bias)
        self.norm_name = None
        if with_norm:
            t = norm_cfg['type']
            if t == 'GN':
                self.norm_name = 'gn'
                self.add_module('gn', nn.GroupNorm(norm_cfg['num_groups'], cout))
            elif t in ('BN', 'SyncBN'):
                self.norm_name = 'bn'
                self.add_module('bn', nn.BatchNorm2d(cout))
            else:
                raise NotImplementedError(f'norm type {t}')
        self.with_act = act

    def forward(self, x):
        x = self.conv(x)
        if self.norm_name is not None:
            x = getattr(self, self.norm_name)(x)
        return F.relu(x) if self.with_act else x


def normal_init_(module, std=0.01, bias=0.0):
    nn.init.normal_(module.weight, 0.0, std)
    if getattr(module, 'bias', None) is not None:
        nn.init.constant_(module.bias, bias)


def bias_init_with_prob(p):
    return float(-math.log((1 - p) / p))


_PACK_ATTRS = ('_ptb_packed_f16', '_ptb_packed_f16_t', '_ptb_packed', '_ptb_packed_tc')


def invalidate_packed(module):
    """drop every cached tensor-core packing of `module`'s weights.  The caches are keyed on (data_ptr, Tensor._version, device);
    in-place writes through `.data` (mmcv's EMAHook swap, manual weight surgery) do NOT bump `_version`, so the heads call this from
    train() / eval() and after load_state_dict, and anything else that writes `.data` must call it explicitly (`.data` writes are
    otherwise unsupported: the packed copy would go stale).  Costs nothing when there is no cache."""
    for m in module.modules():
        for a in _PACK_ATTRS:
            if hasattr(m, a):
                delattr(m, a)


class PackedWeightsMixin:
    """nn.Module mixin of the heads: packed-weight caches are invalidated on train() / eval() and after load_state_dict."""

    def _init_packed_hooks(self):
        self.register_load_state_dict_post_hook(lambda mod, incompatible_keys: invalidate_packed(mod))

    def train(self, mode=True):
        invalidate_packed(self)
        return super().train(mode)

    def invalidate_packed(self):
        invalidate_packed(self)


def _tc_supported(convs, x):
    """the tcgen05 path covers the shipped head geometry: conv3x3 s1 p1 without bias -> 256 channels, GroupNorm with
    channels-per-group % 4 == 0, Cin % 32 == 0, fp32 CUDA input, inference (no autograd graph)."""
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for m in convs for p in m.parameters())):
        return False
    if not x.is_cuda or x.dtype != torch.float32 or len(convs) == 0:
        return False
    for m in convs:
        c = m.conv
        if (c.kernel_size != (3, 3) or c.stride != (1, 1) or c.padding != (1, 1) or c.dilation != (1, 1) or c.groups != 1
                or c.bias is not None or c.out_channels != 256 or c.in_channels % 32 != 0 or m.norm_name != 'gn' or not m.with_act):
            return False
        if (256 // m.gn.num_groups) % 4 != 0 or m.gn.num_groups != 32:
            return False
    return True


def _packed_weight_f16(m):
    from . import ops
    w = m.conv.weight
    key = (w.data_ptr(), w._version, str(w.device), 'f16')
    cache = getattr(m, '_ptb_packed_f16', None)
    if cache is None or cache[0] != key:
        m._ptb_packed_f16 = (key, ops.conv3x3_pack_weight_f16(w))
    return m._ptb_packed_f16[1]


def _packed_weight_f16_t(m):
    """dgrad operand: W^T with the taps reversed ([ci][co][8 - tap]), packed like a forward weight; cached per version."""
    from . import ops
    w = m.conv.weight
    key = (w.data_ptr(), w._version, str(w.device), 'f16t')
    cache = getattr(m, '_ptb_packed_f16_t', None)
    if cache is None or cache[0] != key:
        wt = w.detach().flip(2, 3).transpose(0, 1).reshape(w.shape[1], w.shape[0], 9).contiguous()
        m._ptb_packed_f16_t = (key, ops.conv_tc_pack_weight_f16(wt, 9))
    return m._ptb_packed_f16_t[1]


def _packed_weight(m):
    """TF32 hi/lo packing of a conv weight, cached per parameter version (re-packed after every optimizer step)."""
    from . import ops
    w = m.conv.weight
    key = (w.data_ptr(), w._version, str(w.device))
    cache = getattr(m, '_ptb_packed', None)
    if cache is None or cache[0] != key:
        m._ptb_packed = (key, ops.conv3x3_pack_weight(w))
    return m._ptb_packed[1]


def _packed_tc(module, taps, tag):
    """fp16 (h, l) packing of a Linear / Conv2d weight for ptb_conv_tc_f16x2, cached per parameter version."""
    from . import ops
    w = module.weight
    key = (w.data_ptr(), w._version, str(w.device), tag)
    cache = getattr(module, '_ptb_packed_tc', None)
    if cache is None or cache[0] != key:
        w2 = w.detach().reshape(w.shape[0], w.shape[1], -1) if w.dim() == 4 else w.detach()
        module._ptb_packed_tc = (key, ops.conv_tc_pack_weight_f16(w2.contiguous(), taps))
    return module._ptb_packed_tc[1]


def tc_enabled(x, *modules):
    """inference-only tensor-core path: CUDA fp32, no autograd graph, fp16-split mode selected."""
    import os
    if os.environ.get('PTB_CONV_MODE', 'f16x2') != 'f16x2' or not x.is_cuda or x.dtype != torch.float32:
        return False
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for m in modules for p in m.parameters())):
        return False
    return True


def _tc_train_supported(convs, x):
    """the tensor-core TRAINING tower (forward + dgrad + wgrad + GroupNorm backward kernels) covers the shipped head geometry:
    256 -> 256 conv3x3 s1 p1 without bias, GroupNorm with 8 | channels-per-group, ReLU, fp32 CUDA input."""
    import os
    if os.environ.get('PTB_CONV_MODE', 'f16x2') != 'f16x2' or os.environ.get('PTB_TOWER_TRAIN', 'tc') != 'tc':
        return False
    if not x.is_cuda or x.dtype != torch.float32 or len(convs) == 0:
        return False
    for m in convs:
        c = m.conv
        if not (c.in_channels == 256 and c.out_channels == 256 and c.kernel_size == (3, 3) and c.stride == (1, 1)
                and c.padding == (1, 1) and c.dilation == (1, 1) and c.groups == 1 and c.bias is None and m.norm_name == 'gn'
                and m.with_act and m.gn.num_groups == 32 and m.gn.affine):
            return False
    return True


class _TowerTCFn(torch.autograd.Function):
    """[conv3x3 -> GroupNorm -> ReLU] x n on the tensor cores with a hand-written backward:
    forward  ptb_conv3x3_c256_f16x2 (+ GroupNorm statistics) / ptb_gn_relu_apply[_f16]      (csrc/conv_tc.cu)
    backward ptb_gn_relu_bwd -> ptb_split_f16_amax -> ptb_conv3x3_wgrad_f16x2 (dW) and ptb_conv_tc_f16x2 with the transposed,
             flipped weights (dX)                                                          (csrc/tower_bwd.cu, wgrad_tc.cu)
    Saved per layer: the fp16 operand pair of its input (re-used as the wgrad operand), the conv output y and the statistics."""

    @staticmethod
    def forward(ctx, xm, convs, *params):
        from . import ops
        n_layers = len(convs)
        groups, eps = [m.gn.num_groups for m in convs], [float(m.gn.eps) for m in convs]
        h, l, dev_inv = ops.split_f16(xm, auto_scale=True)
        flag = torch.zeros(1, dtype=torch.int32, device=xm.device)
        saved, out = [], None
        for i in range(n_layers):
            gamma, beta = params[3 * i + 1], params[3 * i + 2]
            wh, wl, inv_w = _packed_weight_f16(convs[i])          # cached per parameter version (no host sync per step)
            inv_x = dev_inv if i == 0 else None
            y, stats = ops.conv3x3_c256_f16(h, l, wh, wl, inv_w, inv_x)
            saved += [h, l, y, stats]
            if i == n_layers - 1:
                out = ops.gn_relu_apply(y, stats, gamma.detach(), beta.detach(), groups[i], eps[i], True, split=False)
            else:
                h, l = ops.gn_relu_apply_f16(y, stats, gamma.detach(), beta.detach(), groups[i], eps[i], True, flag)
        ctx.n_layers, ctx.groups, ctx.eps, ctx.convs = n_layers, groups, eps, convs
        ctx.dev_inv = dev_inv
        ctx.save_for_backward(*saved, *[p.detach() for p in params])
        return out

    @staticmethod
    def backward(ctx, dout):
        from . import ops
        n = ctx.n_layers
        saved = ctx.saved_tensors
        acts, params = saved[:4 * n], saved[4 * n:]
        da = dout.contiguous()
        grads = [None] * (3 * n)
        for i in reversed(range(n)):
            h, l, y, stats = acts[4 * i:4 * i + 4]
            w, gamma, beta = params[3 * i], params[3 * i + 1], params[3 * i + 2]
            dy, dg, db, amax = ops.gn_relu_bwd(da, y, stats, gamma, beta, ctx.groups[i], ctx.eps[i], True)
            dyh, dyl, inv_dy = ops.split_f16_amax(dy, amax)
            grads[3 * i + 1], grads[3 * i + 2] = dg, db
            if ctx.needs_input_grad[2 + 3 * i]:
                grads[3 * i] = ops.conv3x3_wgrad_f16(dyh, dyl, h, l, 1.0, inv_dy, ctx.dev_inv if i == 0 else None)
            if i > 0 or ctx.needs_input_grad[0]:
                # dgrad = the forward kernel with W^T and reversed taps
                da = ops.conv_tc_f16(dyh, dyl, _packed_weight_f16_t(ctx.convs[i]), 9, w.shape[1], dev_out_scale=inv_dy)
            else:
                da = None
        return (da, None, *grads)


def tower(convs, x, info=None, want='fp32'):
    """4 x [conv3x3 + GN + ReLU].  Inference: hand-written tcgen05 implicit GEMM (fp16 two-term split or 3xTF32) with GroupNorm
    statistics in the epilogue (csrc/conv_tc.cu).  Training (autograd): _TowerTCFn for the shipped 256 -> 256 geometry (same forward
    kernel + hand-written backward), else cuDNN fp32 through torch."""
    import os
    mode = os.environ.get('PTB_CONV_MODE', 'f16x2')
    if want == 'f16pair' and not (_tc_supported(convs, x) and mode == 'f16x2' and all(m.conv.in_channels % 32 == 0 for m in convs)):
        return None
    if _tc_supported(convs, x) and mode == 'f16x2' and all(m.conv.in_channels % 32 == 0 for m in convs):
        # two-term fp16 split (22 significant bits), kind::f16: half the tensor-pipe time of 3xTF32.  The first layer's input
        # is scaled by a power of two chosen on the device from max|x| (no host sync); later layers consume GroupNorm outputs.
        from . import ops
        xm = ops.to_nhwc(x).contiguous()
        h, l, dev_inv = ops.split_f16(xm, auto_scale=True)
        flag = torch.zeros(1, dtype=torch.int32, device=x.device)
        out = None
        for i, m in enumerate(convs):
            wh, wl, inv_w = _packed_weight_f16(m)
            y, stats = ops.conv3x3_c256_f16(h, l, wh, wl, inv_w, dev_inv if i == 0 else None)
            if i == len(convs) - 1 and want == 'fp32':
                out = ops.gn_relu_apply(y, stats, m.gn.weight.detach(), m.gn.bias.detach(), m.gn.num_groups, m.gn.eps, True, split=False)
            else:
                h, l = ops.gn_relu_apply_f16(y, stats, m.gn.weight.detach(), m.gn.bias.detach(), m.gn.num_groups, m.gn.eps, True, flag)
        if info is not None:
            info['backend'] = 'tcgen05-f16x2'
            info['overflow_flag'] = flag
        if want == 'f16pair':
            return h, l              # (B,H,W,C) fp16 operand pair of the tower output: feeds ptb_conv_tc_f16x2 directly
        return out.permute(0, 3, 1, 2)
    if _tc_supported(convs, x):
        from . import ops
        xm = ops.to_nhwc(x).contiguous()
        hi, lo = ops.split_tf32(xm)
        out = None
        for i, m in enumerate(convs):
            wh, wl = _packed_weight(m)
            y, stats = ops.conv3x3_c256(hi, lo, wh, wl)
            last = i == len(convs) - 1
            res = ops.gn_relu_apply(y, stats, m.gn.weight.detach(), m.gn.bias.detach(), m.gn.num_groups, m.gn.eps, True,
                                    split=not last)
            if last:
                out = res
            else:
                hi, lo = res
        if info is not None:
            info['backend'] = 'tcgen05-3xtf32'
        return out.permute(0, 3, 1, 2)          # (B,C,H,W) view with channels_last strides
    if want == 'fp32' and torch.is_grad_enabled() and _tc_train_supported(convs, x):
        from . import ops
        params = []
        for m in convs:
            params += [m.conv.weight, m.gn.weight, m.gn.bias]
        out = _TowerTCFn.apply(ops.to_nhwc(x).contiguous(), convs, *params)
        if info is not None:
            info['backend'] = 'tcgen05-f16x2-train'
        return out.permute(0, 3, 1, 2)
    if info is not None:
        info['backend'] = 'cudnn'
    x = x.contiguous(memory_format=torch.channels_last)
    # the head's logits must match the fp32 reference to 1e-4: never let cuDNN drop to TF32 here, whatever the global flag says
    with torch.backends.cudnn.flags(enabled=torch.backends.cudnn.enabled, benchmark=torch.backends.cudnn.benchmark,
                                    deterministic=torch.backends.cudnn.deterministic, allow_tf32=False):
        for m in convs:
            x = m(x)
    return x
