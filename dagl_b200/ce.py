"""Drop-in ``CE`` / ``CES`` modules backed by the sm_100a C-ABI library.

``CE`` mirrors the reference module (DN_Gray/model/dagl.py:174-277): same
constructor signature and defaults, same parameter names and shapes (so a
reference ``state_dict`` / checkpoint loads unchanged, including the unused
``W`` conv), same ``forward(b) -> [B, inter_channels, H, W]``.  The body of
``forward`` is one call through ``dagl_ce_forward_f32`` (include/dagl_b200.h)
(under autograd the backward differentiates a device-side recompute, autograd.py);
there is no PyTorch/CPU fallback — CPU tensors, non-fp32 input, unsupported
configurations or a missing library raise.

``CES`` mirrors dagl.py:74-119 (3 stages x 4 heads, 1x1 merge convs, two
groups of 4 ResBlocks) so the caller row can be exercised without the
reference being importable; its convolutions are plain ``torch.nn`` layers.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from .resblock import ResBlock, fuse_sequential, patch_resblocks

_WS_CACHE: Dict[Tuple[int, int], torch.Tensor] = {}
_WS_LOCK = threading.Lock()


def _workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """Caller-owned workspace, allocated through torch's caching allocator.  One buffer per (device, CUDA stream), grown
    on demand: the kernels of a forward are ordered on the caller's current stream, so two forwards may share scratch
    memory only if they are on the same stream.  Different streams (and the per-device threads of ``nn.DataParallel``, which
    run on different devices) get different buffers; a buffer that is replaced by a larger one is only ever freed to
    torch's allocator on the stream that used it, so the usual stream-ordered reuse rules hold."""
    dev_index = device.index if device.index is not None else torch.cuda.current_device()
    key = (dev_index, torch.cuda.current_stream(device).cuda_stream)
    with _WS_LOCK:
        buf = _WS_CACHE.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(int(nbytes * 1.1) + 1024, dtype=torch.uint8, device=device)
            _WS_CACHE[key] = buf
        return buf


def release_workspaces() -> None:
    """Drop every cached workspace (they are plain torch tensors; memory returns to the caching allocator)."""
    with _WS_LOCK:
        _WS_CACHE.clear()


class CE(nn.Module):
    """Dynamic attentive graph head.  Constructor mirrors dagl.py:175-176."""

    def __init__(self, ksize=7, stride_1=4, stride_2=1, softmax_scale=10, shape=64, p_len=64, in_channels=64,
                 inter_channels=16, use_multiple_size=False, use_topk=False, add_SE=False, num_edge=50,
                 impl: str = "auto", legacy_topk: int = 0):
        super().__init__()
        self.ksize = ksize
        self.shape = shape
        self.p_len = p_len
        self.stride_1 = stride_1
        self.stride_2 = stride_2
        self.softmax_scale = softmax_scale
        self.inter_channels = inter_channels
        self.in_channels = in_channels
        self.use_multiple_size = use_multiple_size
        self.use_topk = use_topk          # stored, never read (as in the reference, dagl.py:187)
        self.add_SE = add_SE
        self.num_edge = num_edge
        self.impl = impl
        # Opt-in extension, NOT reference-compatible behaviour: the shipping reference stores use_topk / num_edge and never
        # reads them, and so does this module.  legacy_topk = k > 0 selects the neighbour rule of the reference's legacy
        # fixed-top-k variant (GReccR2b_3mh_1-checkpoint.py:243-250) instead of the adaptive threshold.
        self.legacy_topk = int(legacy_topk)
        # Parameter containers only: the convolutions/linears below are never
        # *called*; the CUDA kernels read their weights directly.
        self.g = nn.Conv2d(in_channels, inter_channels, kernel_size=3, stride=1, padding=1)
        self.W = nn.Conv2d(inter_channels, in_channels, kernel_size=1, stride=1, padding=0)   # dead in forward
        self.theta = nn.Conv2d(in_channels, inter_channels, kernel_size=1, stride=1, padding=0)
        d = ksize ** 2 * inter_channels
        self.fc1 = nn.Sequential(nn.Linear(d, d // 4), nn.ReLU())
        self.fc2 = nn.Sequential(nn.Linear(d, d // 4), nn.ReLU())
        self.thr_conv = nn.Conv2d(in_channels, 1, kernel_size=ksize, stride=stride_1, padding=0)
        self.bias_conv = nn.Conv2d(in_channels, 1, kernel_size=ksize, stride=stride_1, padding=0)
        self.last_impl: Optional[str] = None
        self.last_launches: int = 0
        self.cache_packed_weights = True      # eval mode: pack fc1/fc2 for the tensor cores once, not per call
        self._packed_buf: Optional[torch.Tensor] = None
        self._packed_key = None

    # -- C-ABI plumbing -------------------------------------------------------
    def _weights(self, device: torch.device) -> Tuple[_lib.DaglCEWeights, list]:
        keep = []

        def ptr(t: torch.Tensor) -> int:
            if t.device != device or t.dtype != torch.float32:
                raise RuntimeError("CE parameters must be fp32 on the input's CUDA device")
            t = t.detach().contiguous()
            keep.append(t)
            return t.data_ptr()

        packed = (self._packed_fc(device) if self.cache_packed_weights and not self.training and self.impl != "simt"
                  else None)
        w = _lib.DaglCEWeights(
            packed_fc=packed,
            g_w=ptr(self.g.weight), g_b=ptr(self.g.bias),
            theta_w=ptr(self.theta.weight), theta_b=ptr(self.theta.bias),
            fc1_w=ptr(self.fc1[0].weight), fc1_b=ptr(self.fc1[0].bias),
            fc2_w=ptr(self.fc2[0].weight), fc2_b=ptr(self.fc2[0].bias),
            thr_w=ptr(self.thr_conv.weight), thr_b=ptr(self.thr_conv.bias),
            bias_w=ptr(self.bias_conv.weight), bias_b=ptr(self.bias_conv.bias),
            in_channels=self.in_channels, inter_channels=self.inter_channels, ksize=self.ksize,
            stride_q=self.stride_1, stride_k=self.stride_2, softmax_scale=float(self.softmax_scale),
            legacy_topk=self.legacy_topk)
        return w, keep

    def _packed_fc(self, device: torch.device) -> Optional[int]:
        """fc1/fc2 and g/theta packed once for the tensor-core kernels (``dagl_ce_pack_weights_f32``) and reused while the
        weights are unchanged (eval mode only; keyed on the tensors' storage and in-place version counters, so
        ``load_state_dict`` / optimiser steps invalidate it)."""
        ws = (self.fc1[0].weight, self.fc2[0].weight, self.g.weight, self.theta.weight, self.fc1[0].bias, self.fc2[0].bias,
              self.g.bias, self.theta.bias)
        key = (str(device),) + tuple(v for t in ws for v in (t.data_ptr(), t._version))
        if self._packed_key != key:
            L = _lib.lib()
            buf = torch.empty(L.dagl_ce_packed_weights_bytes(), dtype=torch.uint8, device=device)
            t1, t2, t3, t4, t5, t6, t7, t8 = (t.detach().contiguous() for t in ws)
            tmp = _lib.DaglCEWeights(fc1_w=t1.data_ptr(), fc2_w=t2.data_ptr(), g_w=t3.data_ptr(), theta_w=t4.data_ptr(),
                                     fc1_b=t5.data_ptr(), fc2_b=t6.data_ptr(), g_b=t7.data_ptr(), theta_b=t8.data_ptr(),
                                     in_channels=self.in_channels, inter_channels=self.inter_channels, ksize=self.ksize)
            rc = L.dagl_ce_pack_weights_f32(C.byref(tmp), buf.data_ptr(), buf.numel(),
                                            torch.cuda.current_stream(device).cuda_stream)
            _lib.check(rc, "dagl_ce_pack_weights_f32")
            self._packed_buf, self._packed_key = buf, key
        return self._packed_buf.data_ptr()

    def _check_input(self, b: torch.Tensor) -> None:
        if not isinstance(b, torch.Tensor) or b.dim() != 4:
            raise RuntimeError("CE.forward expects a [B, C, H, W] tensor")
        if b.dtype != torch.float32:
            raise RuntimeError("CE.forward: fp32 only (the reference default precision)")
        if b.shape[1] != self.in_channels:
            raise RuntimeError(f"CE.forward: expected {self.in_channels} channels, got {b.shape[1]}")

    def _needs_grad(self, b: torch.Tensor) -> bool:
        return torch.is_grad_enabled() and (b.requires_grad or any(p.requires_grad for p in self._grad_params()))

    def _grad_params(self):
        return [self.g.weight, self.g.bias, self.theta.weight, self.theta.bias, self.fc1[0].weight, self.fc1[0].bias,
                self.fc2[0].weight, self.fc2[0].bias, self.thr_conv.weight, self.thr_conv.bias, self.bias_conv.weight,
                self.bias_conv.bias]

    def forward(self, b: torch.Tensor) -> torch.Tensor:
        """``CE.forward`` (dagl.py:207-275).  Under autograd (training, trainer.py:51-57) the value still comes from the
        CUDA path and the backward differentiates a chunked device-side recompute (dagl_b200/autograd.py)."""
        self._check_input(b)
        if not b.is_cuda:
            raise RuntimeError("dagl_b200.CE has no CPU path: input must be a CUDA tensor "
                               "(use forward_host for pinned host buffers)")
        if self._needs_grad(b):
            if self.legacy_topk:
                raise RuntimeError("legacy_topk is an inference-only mode (the backward implements the shipping CE)")
            from .autograd import CEFunction
            return CEFunction.apply(self, b, *self._grad_params())
        return self._forward_cuda(b)

    def _forward_cuda(self, b: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        L = _lib.lib()
        b = b.contiguous()
        B, Cc, H, W = b.shape
        with torch.cuda.device(b.device):
            y = out if out is not None else torch.empty(B, self.inter_channels, H, W, dtype=torch.float32, device=b.device)
            nbytes = L.dagl_ce_workspace_bytes_ex(B, Cc, H, W, _lib.IMPL_BY_NAME[self.impl], 0)
            ws = _workspace(b.device, nbytes)
            w, keep = self._weights(b.device)
            stream = torch.cuda.current_stream(b.device).cuda_stream
            rc = L.dagl_ce_forward_f32(C.byref(w), b.data_ptr(), y.data_ptr(), B, H, W, ws.data_ptr(),
                                       ws.numel(), _lib.IMPL_BY_NAME[self.impl], stream)
            _lib.check(rc, "dagl_ce_forward_f32")
            self.last_impl = L.dagl_last_impl().decode()
            self.last_launches = L.dagl_last_launch_count()
        return y

    def forward_debug(self, b: torch.Tensor):
        """forward + the neighbour selection of dagl.py:256-257.
        Returns (y, mask_bits int32 [B,Nq,ceil(Nk/32)], nnz int32 [B,Nq])."""
        self._check_input(b)
        if not b.is_cuda:
            raise RuntimeError("dagl_b200.CE has no CPU path")
        L = _lib.lib()
        b = b.contiguous()
        B, Cc, H, W = b.shape
        nq = ((H + 3) // 4) * ((W + 3) // 4)
        nw = (H * W + 31) // 32
        with torch.cuda.device(b.device):
            y = torch.empty(B, self.inter_channels, H, W, dtype=torch.float32, device=b.device)
            bits = torch.empty(B, nq, nw, dtype=torch.int32, device=b.device)
            nnz = torch.empty(B, nq, dtype=torch.int32, device=b.device)
            ws = _workspace(b.device, L.dagl_ce_workspace_bytes(B, Cc, H, W))
            w, keep = self._weights(b.device)
            stream = torch.cuda.current_stream(b.device).cuda_stream
            rc = L.dagl_ce_forward_debug_f32(C.byref(w), b.data_ptr(), y.data_ptr(), B, H, W, ws.data_ptr(),
                                             ws.numel(), _lib.IMPL_BY_NAME[self.impl], stream,
                                             bits.data_ptr(), nnz.data_ptr())
            _lib.check(rc, "dagl_ce_forward_debug_f32")
            self.last_impl = L.dagl_last_impl().decode()
            self.last_launches = L.dagl_last_launch_count()
        return y, bits, nnz

    def intermediates(self, b_shape) -> Dict[str, torch.Tensor]:
        """Copies of the prologue results left in the workspace by the last
        forward on this device (parity tests): G, theta, gamma, beta, Q, K, Kbar."""
        L = _lib.lib()
        B, Cc, H, W = b_shape
        dev = self.g.weight.device
        ws = _workspace(dev, L.dagl_ce_workspace_bytes(B, Cc, H, W))
        nq = ((H + 3) // 4) * ((W + 3) // 4)
        nk = H * W
        shapes = {"G": (0, (B, 16, H, W)), "theta": (1, (B, 16, H, W)), "gamma": (2, (B, nq)),
                  "beta": (3, (B, nq)), "Q": (4, (B, nq, 196)), "K": (5, (B, nk, 196)), "Kbar": (6, (B, 196))}
        out = {}
        for name, (which, shp) in shapes.items():
            p = L.dagl_ce_workspace_view(ws.data_ptr(), which, B, Cc, H, W)
            off = (p - ws.data_ptr())
            n = 1
            for d in shp:
                n *= d
            out[name] = ws[off:off + 4 * n].view(torch.float32).view(*shp).clone()
        return out

    def forward_query_sharded(self, b: torch.Tensor, group=None) -> torch.Tensor:
        """Single-image (or small-batch) multi-GPU forward: every rank holds the same input ``b``, runs the cheap
        prologue redundantly and the fused graph stage for its share of the 128-query tiles only; the merged
        aggregation rows are exchanged with ONE all-gather and every rank folds the full result
        (SURVEY §8e scheme 3: rows of the score matrix are independent given all keys).  Without an
        initialised process group this is ``forward``.  Inference only (no autograd graph is recorded)."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return self.forward(b)
        self._check_input(b)
        if not b.is_cuda:
            raise RuntimeError("dagl_b200.CE has no CPU path")
        if self._needs_grad(b):
            raise RuntimeError("forward_query_sharded is an inference path: call it under torch.no_grad() "
                               "(training uses forward(), which records the backward)")
        if self.impl == "simt":
            raise RuntimeError("forward_query_sharded needs a tensor-core impl (auto, tc, tc4)")
        L = _lib.lib()
        b = b.contiguous()
        B, Cc, H, W = b.shape
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        nqt = L.dagl_ce_num_query_tiles(H, W)
        tpr = (nqt + world - 1) // world                      # tiles per rank (last ranks may own fewer / none)
        t0, t1 = min(nqt, rank * tpr), min(nqt, (rank + 1) * tpr)
        rpr = tpr * 128                                       # rows per rank in the exchange buffer
        nq = ((H + 3) // 4) * ((W + 3) // 4)
        with torch.cuda.device(b.device):
            # exchange buffer [world][B][rpr][784]: this rank's kernel writes straight into its slot through a row-offset
            # view (the library indexes rows by global query id), so there is no staging copy on either side
            full = torch.empty(world, B, rpr, 784, dtype=torch.float32, device=b.device)
            stream = torch.cuda.current_stream(b.device).cuda_stream
            if t1 > t0:
                if B != 1:
                    mine = torch.empty(B, nq, 784, dtype=torch.float32, device=b.device)    # the library's row layout: [B][Nq][784]
                    rows_ptr = mine.data_ptr()
                else:
                    mine = None
                    rows_ptr = full[rank].data_ptr() - t0 * 128 * 784 * 4      # row q of the image lives at slot row q - t0*128
                ws = _workspace(b.device, L.dagl_ce_rows_workspace_bytes(B, Cc, H, W, t0, t1))
                w, keep = self._weights(b.device)
                rc = L.dagl_ce_forward_rows_f32(C.byref(w), b.data_ptr(), rows_ptr, B, H, W, t0, t1,
                                                ws.data_ptr(), ws.numel(), _lib.IMPL_BY_NAME[self.impl], stream)
                _lib.check(rc, "dagl_ce_forward_rows_f32")
                self.last_impl = L.dagl_last_impl().decode()
                self.last_launches = L.dagl_last_launch_count()
                if mine is not None:
                    q0, q1 = t0 * 128, min(nq, t1 * 128)
                    full[rank, :, : q1 - q0] = mine[:, q0:q1]
            # in place: rank r's contribution already sits in slot r of the gathered buffer (NCCL's in-place all-gather layout)
            dist.all_gather_into_tensor(full.view(world * B, rpr, 784), full[rank], group=group)
            if B == 1:
                rows = full.view(1, world * rpr, 784)                                    # already in query order
            else:
                rows = full.permute(1, 0, 2, 3).reshape(B, world * rpr, 784)
            if rows.shape[1] != nq or not rows.is_contiguous():
                rows = rows[:, :nq].contiguous()
            y = torch.empty(B, self.inter_channels, H, W, dtype=torch.float32, device=b.device)
            rc = L.dagl_ce_fold_rows_f32(rows.data_ptr(), y.data_ptr(), B, H, W, stream)
            _lib.check(rc, "dagl_ce_fold_rows_f32")
        return y

    def host_pipeline(self, B: int, H: int, W: int, depth: int = 2, device: Optional[torch.device] = None) -> "HostPipeline":
        """Streaming host entry (see ``HostPipeline``): overlaps the host<->device copies of consecutive requests with the
        kernels."""
        return HostPipeline(self, B, H, W, depth, device)

    def forward_host(self, b_host: torch.Tensor, y_host: Optional[torch.Tensor] = None,
                     device: Optional[torch.device] = None, sync: bool = True) -> torch.Tensor:
        """Host-buffer entry (``dagl_ce_forward_host_f32``): ``b_host`` is a CPU
        tensor (pinned for async copies); returns a pinned CPU tensor.  The H2D
        copy, the kernels and the D2H copy are enqueued on the current stream."""
        self._check_input(b_host)
        if b_host.is_cuda:
            raise RuntimeError("forward_host expects a host tensor")
        L = _lib.lib()
        device = device or self.g.weight.device
        b_host = b_host.contiguous()
        B, Cc, H, W = b_host.shape
        if y_host is None:
            y_host = torch.empty(B, self.inter_channels, H, W, dtype=torch.float32).pin_memory()
        elif (not isinstance(y_host, torch.Tensor) or y_host.is_cuda or y_host.dtype != torch.float32 or
              tuple(y_host.shape) != (B, self.inter_channels, H, W) or not y_host.is_contiguous()):
            raise RuntimeError(f"forward_host: y_host must be a contiguous fp32 host tensor of shape "
                               f"{(B, self.inter_channels, H, W)}")
        with torch.cuda.device(device):
            nbytes = L.dagl_ce_workspace_bytes_ex(B, Cc, H, W, _lib.IMPL_BY_NAME[self.impl], 0) + \
                L.dagl_ce_host_staging_bytes(B, Cc, H, W)
            ws = _workspace(device, nbytes)
            w, keep = self._weights(device)
            stream = torch.cuda.current_stream(device).cuda_stream
            rc = L.dagl_ce_forward_host_f32(C.byref(w), b_host.data_ptr(), y_host.data_ptr(), B, H, W,
                                            ws.data_ptr(), ws.numel(), _lib.IMPL_BY_NAME[self.impl], stream)
            _lib.check(rc, "dagl_ce_forward_host_f32")
            self.last_impl = L.dagl_last_impl().decode()
            self.last_launches = L.dagl_last_launch_count()
            if sync:
                torch.cuda.current_stream(device).synchronize()
        return y_host


class HostPipeline:
    """Streaming host entry for a fixed input shape: requests are submitted with host buffers and come back in host buffers,
    and consecutive requests overlap — the H2D copy of request i+1 and the D2H copy of result i-1 run on their own streams
    while the kernels of request i run (``depth`` device slots).  Throughput approaches max(forward, H2D, D2H) per request
    instead of their sum (``CE.forward_host``); the latency of one request is unchanged.

        pipe = ce.host_pipeline(B, H, W)
        for b_host, y_host in requests:        # pinned fp32 host tensors [B,C,H,W] / [B,16,H,W]
            pipe.submit(b_host, y_host)        # asynchronous; returns the event that marks y_host complete
        pipe.drain()
    """

    def __init__(self, ce: "CE", B: int, H: int, W: int, depth: int = 2, device: Optional[torch.device] = None):
        if depth < 2:
            raise ValueError("depth must be >= 2")
        self.ce, self.depth, self.n = ce, depth, 0
        self.device = device or ce.g.weight.device
        if self.device.type != "cuda":
            raise RuntimeError("dagl_b200.CE has no CPU path: move the module to a CUDA device first")
        self.shape_in, self.shape_out = (B, ce.in_channels, H, W), (B, ce.inter_channels, H, W)
        with torch.cuda.device(self.device):
            self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(self.device) for _ in range(3))
            self.b_dev = [torch.empty(self.shape_in, dtype=torch.float32, device=self.device) for _ in range(depth)]
            self.y_dev = [torch.empty(self.shape_out, dtype=torch.float32, device=self.device) for _ in range(depth)]
            self.ev_in = [torch.cuda.Event() for _ in range(depth)]
            self.ev_run = [torch.cuda.Event() for _ in range(depth)]
            self.ev_out = [torch.cuda.Event() for _ in range(depth)]
        cur = torch.cuda.current_stream(self.device)
        for s in (self.s_in, self.s_run, self.s_out):
            s.wait_stream(cur)

    def submit(self, b_host: torch.Tensor, y_host: torch.Tensor) -> "torch.cuda.Event":
        for t, shape, name in ((b_host, self.shape_in, "b_host"), (y_host, self.shape_out, "y_host")):
            if (not isinstance(t, torch.Tensor) or t.is_cuda or t.dtype != torch.float32 or tuple(t.shape) != shape or
                    not t.is_contiguous()):
                raise RuntimeError(f"HostPipeline.submit: {name} must be a contiguous fp32 host tensor of shape {shape}")
        k, used = self.n % self.depth, self.n >= self.depth
        with torch.no_grad():
            with torch.cuda.stream(self.s_in):
                if used:
                    self.s_in.wait_event(self.ev_run[k])         # the previous request of this slot has consumed b_dev[k]
                self.b_dev[k].copy_(b_host, non_blocking=True)
                self.ev_in[k].record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(self.ev_in[k])
                if used:
                    self.s_run.wait_event(self.ev_out[k])        # ... and its result has left y_dev[k]
                self.ce._forward_cuda(self.b_dev[k], out=self.y_dev[k])
                self.ev_run[k].record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_run[k])
                y_host.copy_(self.y_dev[k], non_blocking=True)
                self.ev_out[k].record(self.s_out)
        self.n += 1
        return self.ev_out[k]

    def drain(self) -> None:
        self.s_out.synchronize()
        self.s_run.synchronize()


def stage_heads_forward(heads, x: torch.Tensor) -> torch.Tensor:
    """``torch.cat([h(x) for h in heads], dim=1)`` of dagl.py:114-118 in one C-ABI call
    (``dagl_ces_heads_forward_f32``): every head writes its 16 channels directly into the
    concatenated [B, 16*len(heads), H, W] buffer."""
    heads = list(heads)
    h0 = heads[0]
    for h in heads:
        h._check_input(x)
    if any(h._needs_grad(x) for h in heads):          # training: per-head calls carry the autograd graph
        return torch.cat([h(x) for h in heads], dim=1)
    if not x.is_cuda:
        raise RuntimeError("dagl_b200.CE has no CPU path: input must be a CUDA tensor")
    L = _lib.lib()
    x = x.contiguous()
    B, Cc, H, W = x.shape
    with torch.cuda.device(x.device):
        ycat = torch.empty(B, h0.inter_channels * len(heads), H, W, dtype=torch.float32, device=x.device)
        ws = _workspace(x.device, L.dagl_ces_workspace_bytes(len(heads), B, Cc, H, W))
        structs, keep = [], []
        for h in heads:
            w, k = h._weights(x.device)
            structs.append(w)
            keep.append(k)
        arr = (C.POINTER(_lib.DaglCEWeights) * len(heads))(*[C.pointer(w) for w in structs])
        stream = torch.cuda.current_stream(x.device).cuda_stream
        rc = L.dagl_ces_heads_forward_f32(arr, len(heads), x.data_ptr(), ycat.data_ptr(), B, H, W, ws.data_ptr(),
                                          ws.numel(), _lib.IMPL_BY_NAME[h0.impl], stream)
        _lib.check(rc, "dagl_ces_heads_forward_f32")
        impl = L.dagl_last_impl().decode()
        for h in heads:
            h.last_impl = impl
        h0.last_launches = L.dagl_last_launch_count()
    return ycat


class CES(nn.Module):
    """3 stages x 4 CE heads (dagl.py:74-119); state_dict-compatible with the reference CES."""

    def __init__(self, in_channels: int, num: int = 4, impl: str = "auto"):
        super().__init__()
        self.RBS1 = nn.Sequential(*[ResBlock(in_channels) for _ in range(num)])
        self.RBS2 = nn.Sequential(*[ResBlock(in_channels) for _ in range(num)])
        fuse_sequential(self.RBS1)        # 64 channels: each chain of four is one dagl_resblocks_forward_f32 call
        fuse_sequential(self.RBS2)
        for s in (1, 2, 3):
            for h in (1, 2, 3, 4):
                setattr(self, f"c{s}_{h}", CE(in_channels=in_channels, impl=impl))
            setattr(self, f"c{s}_c", nn.Conv2d(in_channels, in_channels, 1, 1, 0))

    def _stage(self, s: int, x: torch.Tensor) -> torch.Tensor:
        heads = [getattr(self, f"c{s}_{h}") for h in (1, 2, 3, 4)]
        return getattr(self, f"c{s}_c")(stage_heads_forward(heads, x)) + x

    def forward(self, x):
        out = self._stage(1, x)
        out = self.RBS1(out)
        out = self._stage(2, out)
        out = self.RBS2(out)
        return self._stage(3, out)


def _ces_stage_forward(self, x):
    """Body of the reference ``CES.forward`` (dagl.py:112-119) with each ``torch.cat`` of four head calls replaced by one
    stage call (``dagl_ces_heads_forward_f32``: heads as a grid dimension, results written into the concatenated buffer).
    Bound onto the reference's own CES instances by ``patch_reference(..., fuse_stages=True)``; same sub-modules, same
    parameters, bit-identical values."""
    out = self.c1_c(stage_heads_forward((self.c1_1, self.c1_2, self.c1_3, self.c1_4), x)) + x
    out = self.RBS1(out)
    out = self.c2_c(stage_heads_forward((self.c2_1, self.c2_2, self.c2_3, self.c2_4), out)) + out
    out = self.RBS2(out)
    out = self.c3_c(stage_heads_forward((self.c3_1, self.c3_2, self.c3_3, self.c3_4), out)) + out
    return out


_FUSED_CES = {}


def _fused_ces_class(cls):
    """A cached subclass of the reference's own CES class whose ``forward`` makes stage calls.  The instance is switched to
    it in place (same object, sub-modules, state_dict); a class-level override, unlike an instance-bound method, survives
    ``nn.DataParallel``'s module replication (the reference wrapper's multi-GPU path, model/__init__.py:101-103)."""
    if cls not in _FUSED_CES:
        _FUSED_CES[cls] = type(cls.__name__, (cls,), {"forward": _ces_stage_forward, "__module__": cls.__module__,
                                                      "_dagl_fused_stages": True})
    return _FUSED_CES[cls]


def patch_reference(module: nn.Module, impl: str = "auto", fuse_stages: bool = True, fuse_resblocks: bool = True) -> int:
    """Replace every reference ``CE`` instance inside ``module`` (e.g. an ``RR``
    or ``CES`` built by the unmodified reference code) with a ``dagl_b200.CE``
    that *shares* the same Parameters.  Returns the number of heads swapped.

    ``fuse_stages``: also rebind ``forward`` of every reference ``CES`` instance whose twelve heads were swapped, so that
    the four heads of a stage go through ONE stage call instead of four head calls + ``torch.cat`` (same values).

    ``fuse_resblocks``: also route every run of reference ``ResBlock``s inside an ``nn.Sequential`` (``CES.RBS1`` / ``RBS2``,
    ``RR.body``; common.py:59-79) through the tensor-core chain kernel (``dagl_b200.resblock``; fp32-accurate)."""
    n = 0
    for parent in module.modules():
        for name, child in list(parent.named_children()):
            if type(child).__name__ == "CE" and not isinstance(child, CE) and hasattr(child, "thr_conv"):
                new = CE(ksize=child.ksize, stride_1=child.stride_1, stride_2=child.stride_2,
                         softmax_scale=child.softmax_scale, shape=child.shape, p_len=child.p_len,
                         in_channels=child.in_channels, inter_channels=child.inter_channels,
                         use_multiple_size=child.use_multiple_size, use_topk=child.use_topk,
                         add_SE=child.add_SE, num_edge=child.num_edge, impl=impl)
                for sub in ("g", "W", "theta", "fc1", "fc2", "thr_conv", "bias_conv"):
                    setattr(new, sub, getattr(child, sub))
                setattr(parent, name, new)
                n += 1
    if fuse_stages:
        heads = [f"c{s}_{h}" for s in (1, 2, 3) for h in (1, 2, 3, 4)]
        for m in module.modules():
            if type(m).__name__ == "CES" and not isinstance(m, CES) and not getattr(type(m), "_dagl_fused_stages", False) and \
                    all(isinstance(getattr(m, h, None), CE) for h in heads) and hasattr(m, "RBS1") and hasattr(m, "c3_c"):
                m.__class__ = _fused_ces_class(type(m))
    if fuse_resblocks:
        patch_resblocks(module)
    return n


def install(ref_dagl_module, impl: str = "auto") -> None:
    """Rebind ``model.dagl.CE`` so that ``make_model(args)`` of the unmodified
    reference (model/__init__.py:92-93) builds networks on the CUDA head."""
    default_impl = impl

    class _CE(CE):
        def __init__(self, *a, **k):
            k.setdefault("impl", default_impl)
            super().__init__(*a, **k)

    _CE.__name__ = "CE"
    ref_dagl_module.CE = _CE
