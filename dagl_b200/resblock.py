"""ResBlock chains on the tensor-core convolution kernel (``dagl_resblocks_forward_f32``).

Reference: ``common.ResBlock`` (DN_Gray/model/common.py:59-79) -- ``conv3x3 - PReLU - conv3x3``, ``* res_scale``, ``+ x`` --
as used by ``CES.RBS1`` / ``CES.RBS2`` (dagl.py:86-101) and ``RR.body`` (dagl.py:27-34).  A run of consecutive ResBlocks
inside an ``nn.Sequential`` goes through ONE C-ABI call: every convolution is a tcgen05 kernel (split-fp16 x3 operands,
fp32-accurate) whose epilogue applies bias / PReLU / residual and writes the next convolution's operand image.

The fused path is the inference path (no autograd graph): when a gradient is required, or the input is not a CUDA fp32
``[B, 64, H, W]`` tensor, the blocks' own torch modules run (that is the reference's implementation, unchanged).
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import torch
import torch.nn as nn

from . import _lib

MODES = {"pair": 0, "single": 1, "auto": 2}


class ResBlock(nn.Module):
    """Same parameters and names as the reference ResBlock (body.0 / body.1 / body.2; common.py:59-79)."""

    def __init__(self, n_feats: int, res_scale: float = 1.0):
        super().__init__()
        self.body = nn.Sequential(nn.Conv2d(n_feats, n_feats, 3, padding=1), nn.PReLU(),
                                  nn.Conv2d(n_feats, n_feats, 3, padding=1))
        self.res_scale = res_scale

    def forward(self, x):
        if fusable_input(x) and is_resblock(self) and not _needs_grad([self]):
            return resblocks_forward([self], x)
        return self.body(x).mul(self.res_scale) + x


def _conv_ok(c) -> bool:
    return (isinstance(c, nn.Conv2d) and c.in_channels == 64 and c.out_channels == 64 and tuple(c.kernel_size) == (3, 3) and
            tuple(c.stride) == (1, 1) and tuple(c.padding) == (1, 1) and tuple(c.dilation) == (1, 1) and c.groups == 1 and
            c.padding_mode == "zeros" and c.weight.dtype == torch.float32)


def is_resblock(m: nn.Module) -> bool:
    """Duck-typed: the reference's ``ResBlock`` (any task directory) or ours, in the configuration the kernel is built for
    (two 64->64 3x3 convolutions around one PReLU, no batch norm)."""
    body = getattr(m, "body", None)
    if type(m).__name__ != "ResBlock" or not isinstance(body, nn.Sequential) or len(body) != 3:
        return False
    c1, act, c2 = body[0], body[1], body[2]
    return (_conv_ok(c1) and _conv_ok(c2) and isinstance(act, nn.PReLU) and act.weight.numel() in (1, 64) and
            isinstance(getattr(m, "res_scale", 1), (int, float)))


def fusable_input(x: torch.Tensor) -> bool:
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 64):
        return False
    return not (torch.is_grad_enabled() and x.requires_grad)


def _needs_grad(blocks) -> bool:
    return torch.is_grad_enabled() and any(p.requires_grad for b in blocks for p in b.parameters())


def _packed(block: nn.Module, device: torch.device):
    """The block's weights packed for the kernel (``dagl_resblock_pack_weights_f32``), cached on the module while the
    parameters are unchanged (eval mode only)."""
    c1, act, c2 = block.body[0], block.body[1], block.body[2]
    ts = [c1.weight, c1.bias, c2.weight, c2.bias]
    key = tuple((t.data_ptr(), t._version) if t is not None else None for t in ts) + (str(device),)
    cache = block.__dict__.get("_dagl_packed")
    if cache is not None and cache[0] == key:
        return cache[1]
    L = _lib.lib()
    buf = torch.empty(L.dagl_resblock_packed_weights_bytes(), dtype=torch.uint8, device=device)
    w = _struct(block, None)
    rc = L.dagl_resblock_pack_weights_f32(C.byref(w), buf.data_ptr(), buf.numel(), torch.cuda.current_stream(device).cuda_stream)
    _lib.check(rc, "dagl_resblock_pack_weights_f32")
    block.__dict__["_dagl_packed"] = (key, buf)
    return buf


def _struct(block: nn.Module, packed) -> "_lib.DaglResBlockWeights":
    c1, act, c2 = block.body[0], block.body[1], block.body[2]
    ptr = lambda t: None if t is None else t.detach().data_ptr()
    return _lib.DaglResBlockWeights(ptr(c1.weight), ptr(c1.bias), ptr(act.weight), act.weight.numel(), ptr(c2.weight),
                                    ptr(c2.bias), float(getattr(block, "res_scale", 1)),
                                    None if packed is None else packed.data_ptr())


def resblocks_forward(blocks: Sequence[nn.Module], x: torch.Tensor, mode: str = "auto") -> torch.Tensor:
    """``for b in blocks: x = b(x)`` for reference-shaped ResBlocks, in one ``dagl_resblocks_forward_f32`` call."""
    from .ce import _workspace
    blocks = list(blocks)
    if not x.is_cuda:
        raise RuntimeError("dagl_b200 ResBlock chain has no CPU path: input must be a CUDA tensor")
    for b in blocks:
        for t in (b.body[0].weight, b.body[2].weight, b.body[1].weight):
            if not (t.is_cuda and t.is_contiguous() and t.device == x.device):
                raise RuntimeError("ResBlock parameters must be contiguous CUDA tensors on the input's device")
    L = _lib.lib()
    x = x.contiguous()
    B, Cc, H, W = x.shape
    with torch.cuda.device(x.device):
        cache = not any(b.training for b in blocks)
        packed = [_packed(b, x.device) if cache else None for b in blocks]
        arr = (_lib.DaglResBlockWeights * len(blocks))(*[_struct(b, p) for b, p in zip(blocks, packed)])
        y = torch.empty_like(x)
        ws = _workspace(x.device, L.dagl_resblocks_workspace_bytes(len(blocks), B, Cc, H, W))
        rc = L.dagl_resblocks_forward_f32(arr, len(blocks), x.data_ptr(), y.data_ptr(), B, Cc, H, W, ws.data_ptr(),
                                          ws.numel(), MODES[mode], torch.cuda.current_stream(x.device).cuda_stream)
        _lib.check(rc, "dagl_resblocks_forward_f32")
    return y


def _sequential_forward(self, x):
    """``nn.Sequential.forward`` with every run of consecutive ResBlocks taken as one chain call."""
    mods = list(self)
    i = 0
    while i < len(mods):
        if is_resblock(mods[i]) and fusable_input(x):
            j = i
            while j < len(mods) and is_resblock(mods[j]):
                j += 1
            if not _needs_grad(mods[i:j]):
                x = resblocks_forward(mods[i:j], x, getattr(self, "_dagl_resblock_mode", "auto"))
                i = j
                continue
        x = mods[i](x)
        i += 1
    return x


class ChainSequential(nn.Sequential):
    """``nn.Sequential`` whose runs of consecutive ResBlocks are chain calls.  Containers are switched to this class in
    place (``seq.__class__ = ChainSequential``): same object, same sub-modules, same state_dict; being a class-level
    override it also survives ``nn.DataParallel``'s module replication (an instance-bound ``forward`` would not: the
    replica would call the original's sub-modules)."""
    _dagl_resblock_mode = "auto"

    def forward(self, x):
        return _sequential_forward(self, x)


_FUSED_SUBCLASS = {}


def _chain_class(cls):
    """ChainSequential for plain containers, a cached dynamic subclass for user-defined Sequential subclasses."""
    if cls is nn.Sequential:
        return ChainSequential
    if cls not in _FUSED_SUBCLASS:
        _FUSED_SUBCLASS[cls] = type(cls.__name__, (cls,), {"forward": _sequential_forward, "__module__": cls.__module__,
                                                           "_dagl_resblock_mode": "auto", "_dagl_chain": True})
    return _FUSED_SUBCLASS[cls]


def is_fused(seq: nn.Module) -> bool:
    return isinstance(seq, ChainSequential) or getattr(type(seq), "_dagl_chain", False)


def fuse_sequential(seq: nn.Sequential, mode: str = "auto") -> bool:
    """Switch ``seq`` (an ``nn.Sequential`` that holds ResBlocks) to the chained forward, in place.  Same sub-modules and
    parameters; returns whether the container holds anything to fuse."""
    if not isinstance(seq, nn.Sequential) or not any(is_resblock(m) for m in seq):
        return False
    if not is_fused(seq):
        seq.__class__ = _chain_class(type(seq))
    if mode != "auto":
        seq._dagl_resblock_mode = mode
    return True


def patch_resblocks(module: nn.Module, mode: str = "auto") -> int:
    """Fuse the ResBlock runs of every ``nn.Sequential`` inside ``module`` (``CES.RBS1`` / ``RBS2``, ``RR.body``).
    Returns the number of ResBlocks that now run on the chain kernel."""
    n = 0
    for m in module.modules():
        if isinstance(m, nn.Sequential) and fuse_sequential(m, mode):
            n += sum(1 for c in m if is_resblock(c))
    return n
