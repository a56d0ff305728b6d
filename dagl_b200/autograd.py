"""Backward for ``dagl_b200.CE`` (the reference trains through plain autograd,
DN_Gray/trainer.py:51-57; a forward-only head would silently break ``train.py``).

The forward value comes from the CUDA path.  The backward (``CEFunction.backward``) has two parts:

* the graph stage (scores, neighbour mask, softmax, aggregation, fold: dagl.py:250-272) is differentiated by hand-written
  CUDA kernels, ``dagl_graph_attend_backward_f32`` (csrc/attend_bwd.cu): flash-attention style, nothing of the N_q x N_k
  score matrix is kept between forward and backward; it is recomputed per chunk of query rows inside the workspace;
* the plain convolutions / linears in front of it (g, theta, fc1, fc2, thr_conv, bias_conv: dagl.py:208-249) are re-evaluated
  with differentiable torch ops on the device (convolution form of SURVEY.md App. A) and differentiated by autograd, fed
  with the kernel's gradients of Q, K, theta, gamma, beta.

``ce_recompute`` is the all-torch evaluation of the block; it pins the math on the CPU (tests/test_autograd.py) and is
the checker of the CUDA backward, not a fallback of it.

Gradient semantics are the reference's (dagl.py:250-264): the gradient flows through ``S``, through the row mean, through
``mask = relu(S - mu*gamma + beta)`` where it multiplies the logits, and through the softmax; the 0/1 factor ``mask_b``
carries none.  ``W`` (unused in forward) gets no gradient, as in the reference.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn.functional as F
from torch.utils.checkpoint import checkpoint


def _same_pad(n: int, k: int, s: int):
    total = max(0, ((n + s - 1) // s - 1) * s + k - n)
    return total // 2, total - total // 2


def _rows(Qc, Kt, Vt, gamma_c, beta_c, scale: float):
    """Aggregation rows of one query chunk: [n, 784]."""
    S = Qc @ Kt                                                     # [n, Nk]
    mu = S.mean(dim=1, keepdim=True)
    m = F.relu(S - mu * gamma_c.unsqueeze(1) + beta_c.unsqueeze(1))
    P = torch.softmax(S * m * scale, dim=1) * (m != 0).to(S.dtype)
    return P @ Vt


def ce_prologue(b: torch.Tensor, p: Sequence[torch.Tensor], ksize: int = 7, stride_q: int = 4):
    """Differentiable torch evaluation of everything in front of the graph stage (dagl.py:208-249, convolution form):
    returns Q [B,Nq,196], K [B,Nk,196], theta [B,16,H,W], gamma, beta [B,Nq]."""
    g_w, g_b, th_w, th_b, fc1_w, fc1_b, fc2_w, fc2_b, thr_w, thr_b, bias_w, bias_b = p
    B, _, H, W = b.shape
    ci = g_w.shape[0]
    pad_k = ksize // 2
    (pt, pb), (pl, pr) = _same_pad(H, ksize, stride_q), _same_pad(W, ksize, stride_q)
    G = F.conv2d(b, g_w, g_b, padding=1)
    Th = F.conv2d(b, th_w, th_b)
    b4 = F.pad(b, (pl, pr, pt, pb))
    gamma = F.conv2d(b4, thr_w, thr_b, stride=stride_q).flatten(1)
    beta = F.conv2d(b4, bias_w, bias_b, stride=stride_q).flatten(1)
    e = fc1_w.shape[0]
    Q = F.relu(F.conv2d(F.pad(G, (pl, pr, pt, pb)), fc1_w.view(e, ci, ksize, ksize), fc1_b, stride=stride_q))
    K = F.relu(F.conv2d(G, fc2_w.view(e, ci, ksize, ksize), fc2_b, padding=pad_k))
    return (Q.flatten(2).transpose(1, 2).contiguous(), K.flatten(2).transpose(1, 2).contiguous(), Th.contiguous(),
            gamma.contiguous(), beta.contiguous())


def graph_stage_backward(Q, K, Th, gamma, beta, dy, scale: float):
    """Gradients of the graph stage with respect to (Q, K, theta, gamma, beta): one call through the C-ABI
    (``dagl_graph_attend_backward_f32``)."""
    from . import _lib
    from .ce import _workspace
    L = _lib.lib()
    B, _, H, W = Th.shape
    dev = Q.device
    outs = [torch.empty_like(t) for t in (Q, K, Th, gamma, beta)]
    with torch.cuda.device(dev):
        ws = _workspace(dev, L.dagl_graph_attend_backward_workspace_bytes(B, H, W))
        dy = dy.contiguous()
        rc = L.dagl_graph_attend_backward_f32(Q.data_ptr(), K.data_ptr(), Th.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                              dy.data_ptr(), *[o.data_ptr() for o in outs], B, H, W, float(scale),
                                              ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "dagl_graph_attend_backward_f32")
    return outs


def ce_recompute(b: torch.Tensor, p: Sequence[torch.Tensor], ksize: int = 7, stride_q: int = 4, scale: float = 10.0,
                 q_chunk: int = 1024) -> torch.Tensor:
    """Differentiable torch evaluation of the graph block (device tensors).  ``p`` = (g_w, g_b, theta_w, theta_b,
    fc1_w, fc1_b, fc2_w, fc2_b, thr_w, thr_b, bias_w, bias_b)."""
    g_w, g_b, th_w, th_b, fc1_w, fc1_b, fc2_w, fc2_b, thr_w, thr_b, bias_w, bias_b = p
    B, _, H, W = b.shape
    ci = g_w.shape[0]
    pad_k = ksize // 2
    (pt, pb), (pl, pr) = _same_pad(H, ksize, stride_q), _same_pad(W, ksize, stride_q)
    G = F.conv2d(b, g_w, g_b, padding=1)
    Th = F.conv2d(b, th_w, th_b)
    b4 = F.pad(b, (pl, pr, pt, pb))
    gamma = F.conv2d(b4, thr_w, thr_b, stride=stride_q).flatten(1)                       # [B, Nq]
    beta = F.conv2d(b4, bias_w, bias_b, stride=stride_q).flatten(1)
    e = fc1_w.shape[0]
    Q = F.relu(F.conv2d(F.pad(G, (pl, pr, pt, pb)), fc1_w.view(e, ci, ksize, ksize), fc1_b, stride=stride_q))
    K = F.relu(F.conv2d(G, fc2_w.view(e, ci, ksize, ksize), fc2_b, padding=pad_k))
    Q = Q.flatten(2).transpose(1, 2)                                                    # [B, Nq, 196]
    Kt = K.flatten(2)                                                                   # [B, 196, Nk]
    Vt = F.unfold(Th, ksize, padding=pad_k).transpose(1, 2)                             # [B, Nk, 784]
    nq = Q.shape[1]
    outs: List[torch.Tensor] = []
    for i in range(B):
        rows = []
        for s in range(0, nq, q_chunk):
            t = min(nq, s + q_chunk)
            args = (Q[i, s:t], Kt[i], Vt[i], gamma[i, s:t], beta[i, s:t], scale)
            rows.append(checkpoint(_rows, *args, use_reentrant=False) if nq > q_chunk else _rows(*args))
        O = torch.cat(rows, dim=0)                                                      # [Nq, 784]
        outs.append(O.t().unsqueeze(0))
    O = torch.cat(outs, dim=0)                                                          # [B, 784, Nq]
    y = F.fold(O, (H, W), ksize, padding=pad_k, stride=stride_q)
    cnt = F.fold(F.unfold(torch.ones(1, 1, H, W, dtype=b.dtype, device=b.device), ksize, padding=pad_k, stride=stride_q),
                 (H, W), ksize, padding=pad_k, stride=stride_q)
    return y / cnt


class CEFunction(torch.autograd.Function):
    """forward: the CUDA path (``module._forward_cuda``); backward: autograd through ``ce_recompute``."""

    @staticmethod
    def forward(ctx, module, b, *params):
        with torch.no_grad():
            y = module._forward_cuda(b)
        ctx.module = module
        ctx.save_for_backward(b, *params)
        return y

    @staticmethod
    def backward(ctx, dy):
        b, *params = ctx.saved_tensors
        need = ctx.needs_input_grad[1:]
        m = ctx.module
        # The forward's scores are fp32-accurate (3-term split-fp16 MMAs).  The recompute must be too, whatever the training
        # script set globally: TF32 convolutions OR TF32 matmuls (`allow_tf32` / set_float32_matmul_precision) would flip
        # neighbours relative to the forward and put ~1e-3 of error into the gradients.
        # (the two flags are saved / restored by hand: ``torch.backends.cudnn.flags(allow_tf32=False)`` would also switch cuDNN
        # OFF — its ``enabled`` argument defaults to False — and the convolutions below would take the slow native kernels.)
        prev_mm, prev_cd = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        try:
            with torch.enable_grad():
                leaves = [t.detach().requires_grad_(bool(n)) for t, n in zip([b] + params, need)]
                wanted = [t for t in leaves if t.requires_grad]
                # graph stage: CUDA kernels; the convolutions / linears in front of it: torch autograd
                mids = ce_prologue(leaves[0], leaves[1:], ksize=m.ksize, stride_q=m.stride_1)
                with torch.no_grad():
                    dmids = graph_stage_backward(*[t.detach() for t in mids], dy, float(m.softmax_scale))
                grads = list(torch.autograd.grad(list(mids), wanted, dmids, allow_unused=True))
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev_mm
            torch.backends.cudnn.allow_tf32 = prev_cd
        out = [grads.pop(0) if t.requires_grad else None for t in leaves]
        return (None, *out)
