"""``RR`` -- the whole DAGL network (reference: DN_Gray/model/dagl.py:10-54) assembled from this package's modules:
head conv, eight ResBlocks, the ``CES`` module (twelve graph heads), eight ResBlocks, a conv, the tail conv, global skip.
State-dict compatible with the reference's ``RR`` (same sub-module names, incl. the unused ``add_mean`` MeanShift), so the
shipped checkpoints (``*/exp/model/model_best.pt``) load directly.  The ResBlock runs of ``body`` go through the
tensor-core chain kernel (``dagl_b200.resblock``), the heads through the CUDA graph block; the three plain convolutions
(head / body tail / tail: n_colors <-> 64 channels) stay ``nn.Conv2d``.

Task variants: DN_Gray ``RR(n_colors=1)``, CAR ``RR(n_colors=1)`` (its Y-channel model), Demosaic
``RR(n_resblocks=32, n_colors=3)`` (Demosaic/model/dagl.py:14 hard-codes 32 ResBlocks: 16 + CES + 16).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .ce import CES
from .resblock import ResBlock, fuse_sequential


class RR(nn.Module):
    def __init__(self, n_resblocks: int = 16, n_feats: int = 64, n_colors: int = 1, res_scale: float = 1.0,
                 rgb_range: float = 1.0, impl: str = "auto"):
        super().__init__()
        head = nn.Sequential(nn.Conv2d(n_colors, n_feats, 3, padding=1))                                       # dagl.py:24
        body = [ResBlock(n_feats, res_scale) for _ in range(n_resblocks // 2)]                                  # :27-31
        body.append(CES(in_channels=n_feats, impl=impl))                                                        # :22, :32
        body += [ResBlock(n_feats, res_scale) for _ in range(n_resblocks // 2)]                                 # :33-34
        body.append(nn.Conv2d(n_feats, n_feats, 3, padding=1))                                                  # :36
        tail = nn.Sequential(nn.Conv2d(n_feats, n_colors, 3, padding=1))                                       # :37-39
        # common.MeanShift(rgb_range, rgb_mean, rgb_std, +1) (dagl.py:41, common.py:34-45): constructed, never called in forward;
        # registered first, as in the reference, so that the state_dict has the reference's key order too
        self.add_mean = nn.Conv2d(3, 3, 1)
        mean = torch.tensor((0.4488, 0.4371, 0.4040))
        self.add_mean.weight.data = torch.eye(3).view(3, 3, 1, 1)
        self.add_mean.bias.data = rgb_range * mean
        for p in self.add_mean.parameters():
            p.requires_grad = False
        self.head = head
        self.body = nn.Sequential(*body)
        self.tail = tail
        fuse_sequential(self.body)

    def forward(self, x):                                                                                       # dagl.py:47-54
        res = self.head(x)
        res = self.body(res)
        res = self.tail(res)
        return x + res
