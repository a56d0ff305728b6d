"""CPU oracle for DAGL's dynamic attentive graph block (``CE`` / ``CES``).

TEST INFRASTRUCTURE ONLY.  Nothing under ``dagl_b200/`` may import this module.
It is the checker used by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``; it is never the
thing shipped and it is never a fallback for the CUDA path.

It is a restatement (not a copy) of the reference algorithm in plain PyTorch
CPU fp32/fp64, following

* ``DN_Gray/model/dagl.py:123-139``  (``same_padding``: TF "SAME" pad rule)
* ``DN_Gray/model/dagl.py:142-169``  (``extract_image_patches``: pad + Unfold)
* ``DN_Gray/model/dagl.py:207-275``  (``CE.forward``)
* ``DN_Gray/model/dagl.py:112-119``  (``CES.forward``)
* ``DN_Gray/model/common.py:59-79``  (``ResBlock``), used inside ``CES``

Parity pin: ``oracle/make_golden.py`` imports the *unmodified* reference from
``/root/reference`` in the build container, runs it on seeded inputs and
stores inputs/weights/outputs/neighbour masks under ``tests/golden/``;
``tests/test_oracle.py`` checks every function here against those fixtures
(bit-exact for ``ce_forward`` in reference order).  The reference ships no
golden vectors of its own for this path.

Two evaluation orders are provided:

``ce_forward``            op-for-op in the reference's order (materialises the
                          N_q x N_k score matrix; needs ~20 x 4 N_q N_k bytes).
``ce_forward_chunked``    same math, query rows processed in chunks so 256^2
                          and 512^2 inputs fit in host memory; rows are
                          independent so results are identical up to the BLAS
                          blocking of the row chunk.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

KSIZE = 7
STRIDE_Q = 4
STRIDE_K = 1
SOFTMAX_SCALE = 10.0

CE_PARAM_SHAPES = {
    "g.weight": (16, 64, 3, 3), "g.bias": (16,),
    "W.weight": (64, 16, 1, 1), "W.bias": (64,),          # present, unused in forward
    "theta.weight": (16, 64, 1, 1), "theta.bias": (16,),
    "fc1.0.weight": (196, 784), "fc1.0.bias": (196,),
    "fc2.0.weight": (196, 784), "fc2.0.bias": (196,),
    "thr_conv.weight": (1, 64, 7, 7), "thr_conv.bias": (1,),
    "bias_conv.weight": (1, 64, 7, 7), "bias_conv.bias": (1,),
}


def same_pad_amounts(n: int, k: int, s: int) -> Tuple[int, int]:
    """(before, after) zero padding of TF-"SAME" for one dimension.

    dagl.py:126-136 — total = max(0, (ceil(n/s)-1)*s + k - n); the smaller half
    goes first (top/left), the remainder last (bottom/right)."""
    out = (n + s - 1) // s
    total = max(0, (out - 1) * s + k - n)
    before = total // 2
    return before, total - before


def num_queries(H: int, W: int) -> Tuple[int, int]:
    return (H + STRIDE_Q - 1) // STRIDE_Q, (W + STRIDE_Q - 1) // STRIDE_Q


def ce_param_shapes(in_channels: int = 64):
    shapes = dict(CE_PARAM_SHAPES)
    for name in ("g.weight", "theta.weight", "thr_conv.weight", "bias_conv.weight"):
        s = list(shapes[name]); s[1] = in_channels; shapes[name] = tuple(s)
    shapes["W.weight"] = (in_channels, 16, 1, 1)
    shapes["W.bias"] = (in_channels,)
    return shapes


def init_ce_params(seed: int, in_channels: int = 64, dtype=torch.float32) -> Params:
    """Random CE parameters with torch's default Conv2d/Linear init statistics
    (weight and bias both U(-1/sqrt(fan_in), 1/sqrt(fan_in))).  Used where a
    reference-constructed module is not available (GPU box); the golden
    fixtures carry reference-made weights."""
    gen = torch.Generator().manual_seed(seed)
    shapes = ce_param_shapes(in_channels)
    p: Params = {}
    for name, shape in shapes.items():
        wshape = shapes[name.rsplit(".", 1)[0] + ".weight"]
        fan_in = 1
        for d in wshape[1:]:
            fan_in *= d
        bound = 1.0 / math.sqrt(fan_in)
        p[name] = ((torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
    return p


def _prologue(p: Params, b: torch.Tensor):
    """dagl.py:208-243 — feature maps, per-query threshold/bias, patch tensors."""
    B, C, H, W = b.shape
    G = F.conv2d(b, p["g.weight"], p["g.bias"], padding=1)            # :208  query AND key source (:210)
    Th = F.conv2d(b, p["theta.weight"], p["theta.bias"])              # :209  value source
    pt, pb = same_pad_amounts(H, KSIZE, STRIDE_Q)
    pl, pr = same_pad_amounts(W, KSIZE, STRIDE_Q)
    b4 = F.pad(b, (pl, pr, pt, pb))                                   # :213
    gamma = F.conv2d(b4, p["thr_conv.weight"], p["thr_conv.bias"], stride=STRIDE_Q).reshape(B, -1)   # :214
    beta = F.conv2d(b4, p["bias_conv.weight"], p["bias_conv.bias"], stride=STRIDE_Q).reshape(B, -1)  # :215
    qp = F.unfold(F.pad(G, (pl, pr, pt, pb)), KSIZE, stride=STRIDE_Q)  # :216-221  [B,784,Nq]
    kt, kb = same_pad_amounts(H, KSIZE, STRIDE_K)
    kl, kr = same_pad_amounts(W, KSIZE, STRIDE_K)
    vp = F.unfold(F.pad(Th, (kl, kr, kt, kb)), KSIZE, stride=STRIDE_K)  # :224-230 [B,784,Nk]
    kp = F.unfold(F.pad(G, (kl, kr, kt, kb)), KSIZE, stride=STRIDE_K)   # :233-239 [B,784,Nk]
    fold_pad = kl                                                      # :243,267 — only paddings[0] is used
    return G, Th, gamma, beta, qp, kp, vp, fold_pad


def _fold_normalise(O: torch.Tensor, H: int, W: int, fold_pad: int) -> torch.Tensor:
    """dagl.py:265-272 — overlap-add of the 7x7x16 output patches, divided by
    the coverage count (a constant of (H, W))."""
    Nq = O.shape[0]
    zi = O.reshape(1, Nq, -1).permute(0, 2, 1)
    zi = F.fold(zi, (H, W), (KSIZE, KSIZE), padding=fold_pad, stride=STRIDE_Q)
    ones = torch.ones_like(zi)
    cnt = F.fold(F.unfold(ones, (KSIZE, KSIZE), padding=fold_pad, stride=STRIDE_Q),
                 (H, W), (KSIZE, KSIZE), padding=fold_pad, stride=STRIDE_Q)
    cnt = cnt + (cnt == 0.).to(cnt.dtype)                              # :271 (no-op: cnt >= 1)
    return zi / cnt


def _edge_weights(S: torch.Tensor, mu: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor):
    """dagl.py:256-261 — adaptive neighbour mask and masked, un-renormalised softmax.

    S [n, Nk]; mu, gamma, beta [n]."""
    mask = F.relu(S - mu.unsqueeze(1) * gamma.unsqueeze(1) + beta.unsqueeze(1))   # :256
    mask_b = (mask != 0.).to(S.dtype)                                              # :257
    P = F.softmax((S * mask) * SOFTMAX_SCALE, dim=1)                               # :259-260
    return P * mask_b, mask_b                                                      # :261


def ce_forward(p: Params, b: torch.Tensor, return_aux: bool = False):
    """Reference-order CE forward.  b [B,C,H,W] -> y [B,16,H,W].

    With ``return_aux`` also returns dict(mask=[B,Nq,Nk] bool, Q, K, gamma, beta,
    mu, S) for the parity tests."""
    B, C, H, W = b.shape
    G, Th, gamma, beta, qp, kp, vp, fold_pad = _prologue(p, b)
    ys, aux = [], dict(mask=[], Q=[], K=[], mu=[], S=[])
    for i in range(B):                                                            # :245
        Q = F.relu(F.linear(qp[i].t(), p["fc1.0.weight"], p["fc1.0.bias"]))      # :248 [Nq,196]
        K = F.relu(F.linear(kp[i].t(), p["fc2.0.weight"], p["fc2.0.bias"]))      # :249 [Nk,196]
        S = torch.matmul(Q, K.t())                                                # :250
        mu = S.mean(dim=1)                                                        # :256
        P, mask_b = _edge_weights(S, mu, gamma[i], beta[i])
        O = torch.mm(P, vp[i].t())                                                # :263-264 [Nq,784]
        ys.append(_fold_normalise(O, H, W, fold_pad))
        if return_aux:
            aux["mask"].append(mask_b.bool()); aux["Q"].append(Q); aux["K"].append(K)
            aux["mu"].append(mu); aux["S"].append(S)
    y = torch.cat(ys, dim=0)                                                      # :274
    if not return_aux:
        return y
    out = {k: torch.stack(v) for k, v in aux.items()}
    out.update(gamma=gamma, beta=beta, G=G, theta=Th)
    return y, out


def ce_forward_chunked(p: Params, b: torch.Tensor, chunk: int = 512, return_nnz: bool = False,
                       other_mask_bits: Optional[torch.Tensor] = None):
    """Same math with query rows processed ``chunk`` at a time (rows of S are
    independent: dagl.py:250-264 are all row-wise).  Memory ~ 6 * chunk * Nk
    floats instead of ~20 * Nq * Nk.

    ``other_mask_bits`` ([B, Nq, ceil(Nk/32)] int32, the layout of ``dagl_ce_forward_debug_f32``): compare another
    implementation's neighbour mask with the oracle's, chunk by chunk, and return as a third result the list of flipped
    entries ``(image, query, key, margin_in_ulps)`` with margin = |S - mu*gamma + beta| / (eps * (|S| + |mu*gamma| + |beta|))
    — how far the flipped entry sits from the relu threshold of dagl.py:256 in units of the fp32 rounding of its operands."""
    B, C, H, W = b.shape
    G, Th, gamma, beta, qp, kp, vp, fold_pad = _prologue(p, b)
    ys, nnzs, flips = [], [], []
    eps = torch.finfo(torch.float32).eps
    for i in range(B):
        Q = F.relu(F.linear(qp[i].t(), p["fc1.0.weight"], p["fc1.0.bias"]))
        K = F.relu(F.linear(kp[i].t(), p["fc2.0.weight"], p["fc2.0.bias"]))
        Kt = K.t().contiguous()
        V = vp[i].t().contiguous()
        Nq = Q.shape[0]
        O = torch.empty(Nq, V.shape[1], dtype=b.dtype)
        nnz = torch.empty(Nq, dtype=torch.int64)
        for s in range(0, Nq, chunk):
            e = min(Nq, s + chunk)
            S = torch.matmul(Q[s:e], Kt)
            mu = S.mean(dim=1)
            P, mask_b = _edge_weights(S, mu, gamma[i, s:e], beta[i, s:e])
            nnz[s:e] = mask_b.sum(dim=1).to(torch.int64)
            O[s:e] = torch.mm(P, V)
            if other_mask_bits is not None:
                diff = unpack_mask_bits(other_mask_bits[i, s:e], S.shape[1]) != mask_b.bool()
                if bool(diff.any()):
                    t = (mu * gamma[i, s:e]).unsqueeze(1)
                    bt = beta[i, s:e].unsqueeze(1)
                    margin = ((S - t) + bt).abs() / (eps * (S.abs() + t.abs() + bt.abs()))
                    for r, k in diff.nonzero().tolist():
                        flips.append((i, s + r, k, float(margin[r, k])))
        ys.append(_fold_normalise(O, H, W, fold_pad))
        nnzs.append(nnz)
    y = torch.cat(ys, dim=0)
    if other_mask_bits is not None:
        return y, torch.stack(nnzs), flips
    return (y, torch.stack(nnzs)) if return_nnz else y


def ce_forward_topk(p: Params, b: torch.Tensor, k: int, return_mask: bool = False):
    """The graph stage of the reference's LEGACY fixed-top-k variant
    (DN_Gray/model/.ipynb_checkpoints/GReccR2b_3mh_1-checkpoint.py:196-262; no entry point of the reference imports it):
    same patches / embeddings / fold as the shipping CE, but the neighbours of a query are its ``min(k, Nk)`` highest-scoring
    keys (:243-247), the logits are ``scale * S`` on them and 0 elsewhere (:248-249), and there is no thr / bias branch.
    Returns the folded, count-normalised aggregation (``zi / out_mask``, :255-260), i.e. without that variant's trailing
    ``W`` conv and residual (:262-263), which the shipping ``CES`` applies outside ``CE``."""
    B, C, H, W = b.shape
    G, Th, _, _, qp, kp, vp, fold_pad = _prologue(p, b)
    ys, masks = [], []
    for i in range(B):
        Q = F.relu(F.linear(qp[i].t(), p["fc1.0.weight"], p["fc1.0.bias"]))      # :236
        K = F.relu(F.linear(kp[i].t(), p["fc2.0.weight"], p["fc2.0.bias"]))      # :237
        S = torch.matmul(Q, K.t())                                                # :238
        top_k = min(k, S.shape[1])                                                # :243
        _, pred = torch.topk(S, top_k, dim=1)                                     # :244
        mask = torch.zeros_like(S)
        mask.scatter_(1, pred, 1.0)                                               # :245-247
        yi = F.softmax((S * mask) * SOFTMAX_SCALE, dim=1) * mask                  # :248-250
        O = torch.mm(yi, vp[i].t())                                               # :253
        ys.append(_fold_normalise(O, H, W, fold_pad))                             # :255-260
        masks.append(mask.bool())
    y = torch.cat(ys, dim=0)
    return (y, torch.stack(masks)) if return_mask else y


# ---------------------------------------------------------------------------
# Caller row: CES = 3 stages x 4 heads + ResBlocks
# ---------------------------------------------------------------------------

def split_ces_state(state: Params, prefix: str = "") -> Dict[str, Params]:
    """Group a CES state_dict by sub-module name (c1_1 ... c3_4, c1_c, RBS1.0 ...)."""
    out: Dict[str, Params] = {}
    for k, v in state.items():
        if prefix and not k.startswith(prefix):
            continue
        k2 = k[len(prefix):]
        head, rest = k2.split(".", 1)
        if head in ("RBS1", "RBS2"):
            idx, rest = rest.split(".", 1)
            head = f"{head}.{idx}"
        out.setdefault(head, {})[rest] = v
    return out


def _resblock(p: Params, x: torch.Tensor) -> torch.Tensor:
    """common.py:59-79 with act=PReLU, res_scale=1: conv3x3 - PReLU - conv3x3, + x."""
    r = F.conv2d(x, p["body.0.weight"], p["body.0.bias"], padding=1)
    r = F.prelu(r, p["body.1.weight"])
    r = F.conv2d(r, p["body.2.weight"], p["body.2.bias"], padding=1)
    return r + x


def ces_forward(state: Params, x: torch.Tensor, ce_fn=ce_forward) -> torch.Tensor:
    """dagl.py:112-119.  ``ce_fn(params, x)`` evaluates one head (defaults to the
    oracle; tests pass the CUDA module's forward to check the caller row)."""
    mods = split_ces_state(state)
    out = x
    for stage in (1, 2, 3):
        heads = [ce_fn(mods[f"c{stage}_{h}"], out) for h in (1, 2, 3, 4)]
        cat = torch.cat(heads, dim=1)
        out = F.conv2d(cat, mods[f"c{stage}_c"]["weight"], mods[f"c{stage}_c"]["bias"]) + out
        if stage < 3:
            for r in range(4):
                out = _resblock(mods[f"RBS{stage}.{r}"], out)
    return out


# ---------------------------------------------------------------------------
# helpers shared by tests / bench
# ---------------------------------------------------------------------------

def pack_mask_bits(mask: torch.Tensor) -> torch.Tensor:
    """[..., Nk] bool -> [..., ceil(Nk/32)] int32, bit j of word w = key 32w+j
    (the layout ``dagl_ce_forward_debug_f32`` writes)."""
    Nk = mask.shape[-1]
    nw = (Nk + 31) // 32
    m = F.pad(mask.to(torch.int64), (0, nw * 32 - Nk)).reshape(*mask.shape[:-1], nw, 32)
    w = (m << torch.arange(32, dtype=torch.int64)).sum(dim=-1)
    w = torch.where(w >= 2 ** 31, w - 2 ** 32, w)
    return w.to(torch.int32)


def unpack_mask_bits(words: torch.Tensor, Nk: int) -> torch.Tensor:
    w = words.to(torch.int64) & 0xFFFFFFFF
    bits = (w.unsqueeze(-1) >> torch.arange(32, dtype=torch.int64)) & 1
    return bits.reshape(*words.shape[:-1], -1)[..., :Nk].bool()


def algorithmic_flops(B: int, H: int, W: int) -> float:
    """2 * Nq * Nk * (E + D) per image (dense)."""
    nqy, nqx = num_queries(H, W)
    return 2.0 * B * nqy * nqx * H * W * (196 + 784)


def algorithmic_bytes(B: int, H: int, W: int) -> float:
    """fp32 bytes the fused graph kernel must move per image:
    4 * [Nq*E + Nk*E + E + 2*Nq + Ci*HW (theta in) + Ci*HW (y out)]."""
    nqy, nqx = num_queries(H, W)
    nq, nk = nqy * nqx, H * W
    return 4.0 * B * (nq * 196 + nk * 196 + 196 + 2 * nq + 16 * nk + 16 * nk)
