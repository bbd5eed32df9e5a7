"""Host-side layer containers with the reference's parameter names (checkpoint compatibility).

The conv towers are library GEMMs (cuDNN through torch, channels_last); they are NOT part of the hand-written hot path
yet (tcgen05 implicit-GEMM towers are the next row) — see DESIGN.md.
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


def tower(convs, x):
    x = x.contiguous(memory_format=torch.channels_last)
    for m in convs:
        x = m(x)
    return x
