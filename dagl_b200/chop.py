"""Tile scheduler for inference: the reference's ``forward_chop`` / x8 self-ensemble, batched.

Reference: ``Model.forward_chop`` (DN_Gray/model/__init__.py:179-231) quarters the image recursively
into four overlapping tiles until a tile has fewer than ``min_size`` pixels, runs the network on every
leaf tile and stitches the quadrant centres back; ``test_x8`` (:53-62, helpers :18-51) averages the
network over the 8 flips/rotations of its input, going through numpy on the host for every transform.

Here the recursion is *planned* up front (``plan``): every leaf has the same size, so all leaves of all
images go through the network as ONE batch dimension (chunked by ``tile_batch``), the flips/rotations
are torch ops on the device, and with an initialised process group the leaf tiles are sharded over the
ranks with no data-path collective (SURVEY.md §8e scheme 2; one all-gather of the tile outputs).
Results are identical to the reference's because the same pixels of the same tiles are computed and
copied; only the order of the launches changes.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Tuple

import torch

from . import parallel


@dataclass(frozen=True)
class LeafCopy:
    """One leaf tile and the part of it that survives the stitching."""
    tile: Tuple[int, int, int, int]     # (y0, x0, h, w) of the tile in the input image
    src: Tuple[int, int, int, int]      # (y0, y1, x0, x1) inside the tile
    dst: Tuple[int, int, int, int]      # (y0, y1, x0, x1) in the output image


def _quadrants(h: int, w: int, shave_size_max: int, shave_scale: int):
    """Tile size and the four (tile origin, kept window) pairs of one recursion level
    (model/__init__.py:185-199, 216-229)."""
    h_half, w_half = h // 2, w // 2
    h_size = (h_half // shave_scale) * shave_scale + shave_size_max
    w_size = (w_half // shave_scale) * shave_scale + shave_size_max
    quads = [
        # tile origin (y, x)      kept rows in tile              kept cols in tile                output origin
        ((0, 0),                 (0, h_half),                   (0, w_half),                     (0, 0)),
        ((0, w - w_size),        (0, h_half),                   (w_size - w + w_half, w_size),   (0, w_half)),
        ((h - h_size, 0),        (h_size - h + h_half, h_size), (0, w_half),                     (h_half, 0)),
        ((h - h_size, w - w_size), (h_size - h + h_half, h_size), (w_size - w + w_half, w_size), (h_half, w_half)),
    ]
    return h_size, w_size, quads


def plan(h: int, w: int, shave_size_max: int = 24, shave_scale: int = 4, min_size: int = 10000) -> List[LeafCopy]:
    """Leaf tiles of ``forward_chop`` for an h x w image and what each contributes to the output.
    ``shave_size_max`` is 24 in DN_Gray / CAR, 12 in Demosaic / DN_Real."""
    if h <= 0 or w <= 0:
        raise ValueError("empty image")
    out: List[LeafCopy] = []

    def rec(y0: int, x0: int, hh: int, ww: int, keep: Tuple[int, int, int, int], oy: int, ox: int):
        # node = image region (y0, x0, hh, ww); `keep` = window of the node's output (node coordinates)
        # that reaches the final image, whose top-left lands at (oy, ox) of the final output.
        h_size, w_size, quads = _quadrants(hh, ww, shave_size_max, shave_scale)
        if h_size > hh or w_size > ww or h_size <= 0 or w_size <= 0:
            raise ValueError(f"image region {hh}x{ww} is too small to chop with shave {shave_size_max}")
        leaf = w_size * h_size < min_size
        for (ty, tx), (ry0, ry1), (rx0, rx1), (ny, nx) in quads:
            # region of the node's output filled by this quadrant: rows [ny, ny + ry1-ry0), cols [nx, ...)
            ky0, ky1 = max(keep[0], ny), min(keep[1], ny + (ry1 - ry0))
            kx0, kx1 = max(keep[2], nx), min(keep[3], nx + (rx1 - rx0))
            if ky0 >= ky1 or kx0 >= kx1:
                continue
            # the same window in tile coordinates
            sy0, sy1 = ky0 - ny + ry0, ky1 - ny + ry0
            sx0, sx1 = kx0 - nx + rx0, kx1 - nx + rx0
            d_y, d_x = oy + (ky0 - keep[0]), ox + (kx0 - keep[2])
            if leaf:
                out.append(LeafCopy((y0 + ty, x0 + tx, h_size, w_size), (sy0, sy1, sx0, sx1),
                                    (d_y, d_y + (sy1 - sy0), d_x, d_x + (sx1 - sx0))))
            else:
                rec(y0 + ty, x0 + tx, h_size, w_size, (sy0, sy1, sx0, sx1), d_y, d_x)

    rec(0, 0, h, w, (0, h, 0, w), 0, 0)
    return out


# ---- x8 self-ensemble (model/__init__.py:18-62) on the device ------------------------------------
def augment(x: torch.Tensor, mode: int) -> torch.Tensor:
    """The reference's ``augment_img`` on the (H, W) axes of a [B, C, H, W] tensor, as torch ops
    (np.rot90 on the first two axes of an (H, W, ..) array == torch.rot90 over dims (2, 3);
    np.flipud == flip of H)."""
    r = lambda t, k: torch.rot90(t, k, dims=(2, 3))
    f = lambda t: torch.flip(t, dims=(2,))
    if mode == 0:
        return x
    if mode == 1:
        return f(r(x, 1))
    if mode == 2:
        return f(x)
    if mode == 3:
        return r(x, 3)
    if mode == 4:
        return f(r(x, 2))
    if mode == 5:
        return r(x, 1)
    if mode == 6:
        return r(x, 2)
    if mode == 7:
        return f(r(x, 3))
    raise ValueError("mode must be 0..7")


def forward_x8(model: Callable[[torch.Tensor], torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """Mean of the model over the 8 flips/rotations (``test_x8``); modes 3 and 5 are each other's inverse."""
    outs = []
    for i in range(8):
        y = model(augment(x, i).contiguous())
        outs.append(augment(y, 8 - i if i in (3, 5) else i))
    return torch.stack(outs, dim=0).mean(dim=0)


def forward_x8_flips(forward_function: Callable[[torch.Tensor], torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """``Model.forward_x8`` of the Demosaic / DN_Real wrappers (Demosaic/model/__init__.py:265-299), the ``self_ensemble``
    branch of ``Model.forward`` (:107-114): the 8 inputs are built by successively appending the W-flip ('v'), H-flip ('h')
    and transpose ('t') of everything so far; output i is un-transposed if i > 3, un-H-flipped if i % 4 > 1, un-W-flipped
    if i is odd; the result is the mean over the concatenated batch axis (so, as in the reference, one image per call).
    ``forward_function`` is the network or ``lambda t: forward_chop(net, t, shave_size_max=12)``."""
    tf = {"v": lambda t: torch.flip(t, dims=(3,)), "h": lambda t: torch.flip(t, dims=(2,)),
          "t": lambda t: t.transpose(2, 3)}
    lr = [x]
    for op in ("v", "h", "t"):
        lr.extend([tf[op](t).contiguous() for t in lr])
    sr = [forward_function(t) for t in lr]
    for i in range(len(sr)):
        if i > 3:
            sr[i] = tf["t"](sr[i])
        if i % 4 > 1:
            sr[i] = tf["h"](sr[i])
        if (i % 4) % 2 == 1:
            sr[i] = tf["v"](sr[i])
    return torch.cat(sr, dim=0).mean(dim=0, keepdim=True)


# ---- the scheduler ---------------------------------------------------------------------------------
def forward_chop(model: Callable[[torch.Tensor], torch.Tensor], x: torch.Tensor, ensemble: bool = False,
                 shave_size_max: int = 24, shave_scale: int = 4, min_size: int = 10000,
                 tile_batch: Optional[int] = 64, distributed: bool = True) -> torch.Tensor:
    """``Model.forward_chop`` (scale 1) with all leaf tiles batched through ``model``.

    ``model`` maps [n, C, th, tw] -> [n, C', th, tw] (e.g. a reference ``RR`` whose ``CE`` heads were
    swapped by ``dagl_b200.patch_reference``).  ``tile_batch`` bounds the tiles per network call.
    With an initialised ``torch.distributed`` group (and ``distributed=True``) each rank runs its
    contiguous share of the tiles; the tile outputs are all-gathered and every rank stitches."""
    if x.dim() != 4:
        raise ValueError("forward_chop expects [B, C, H, W]")
    b, _, h, w = x.shape
    leaves = plan(h, w, shave_size_max, shave_scale, min_size)
    # distinct tiles in first-use order (a tile appears once per recursion leaf)
    tiles: List[Tuple[int, int, int, int]] = []
    index = {}
    for lc in leaves:
        if lc.tile not in index:
            index[lc.tile] = len(tiles)
            tiles.append(lc.tile)
    batch = torch.cat([x[:, :, ty:ty + th, tx:tx + tw] for (ty, tx, th, tw) in tiles], dim=0).contiguous()

    run = (lambda t: forward_x8(model, t)) if ensemble else model

    def run_chunked(t: torch.Tensor) -> torch.Tensor:
        if not tile_batch or t.shape[0] <= tile_batch:
            return run(t)
        return torch.cat([run(t[i:i + tile_batch]) for i in range(0, t.shape[0], tile_batch)], dim=0)

    with torch.no_grad():
        outs = parallel.forward_sharded(run_chunked, batch, gather=True) if distributed else run_chunked(batch)
    out = None
    for lc in leaves:
        t = index[lc.tile]
        tile_out = outs[t * b:(t + 1) * b]
        if out is None:
            out = torch.empty((b, tile_out.shape[1], h, w), dtype=tile_out.dtype, device=tile_out.device)
        sy0, sy1, sx0, sx1 = lc.src
        dy0, dy1, dx0, dx1 = lc.dst
        out[:, :, dy0:dy1, dx0:dx1] = tile_out[:, :, sy0:sy1, sx0:sx1]
    return out
