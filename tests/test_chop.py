"""Tile scheduler (dagl_b200/chop.py) against fixtures produced by the UNMODIFIED reference wrapper
(``Model.forward_chop`` / ``test_x8``, DN_Gray/model/__init__.py; fixtures: oracle/make_golden_chop.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_npz
from dagl_b200 import chop


def probe_net(t):
    """Position-dependent stand-in network (identical to oracle/make_golden_chop.py::probe_net)."""
    n, c, h, w = t.shape
    yy = torch.arange(h, dtype=t.dtype, device=t.device).view(1, 1, h, 1)
    xx = torch.arange(w, dtype=t.dtype, device=t.device).view(1, 1, 1, w)
    return t * (1.0 + 0.01 * yy + 0.0003 * xx) + 0.125 * yy - 0.0625 * xx


def test_plan_tiles_match_survey_appendix_d():
    # leaves per image and their size
    cases = [((256, 256, 24), 64, (72, 72)), ((512, 512, 24), 256, (76, 76)), ((321, 481, 24), 64, (80, 100)),
             ((256, 256, 12), 16, (80, 80)), ((512, 512, 12), 64, (84, 84))]
    for (h, w, shave), n, size in cases:
        leaves = chop.plan(h, w, shave_size_max=shave)
        tiles = {lc.tile for lc in leaves}
        assert len(tiles) == n, (h, w, shave, len(tiles))
        assert {(t[2], t[3]) for t in tiles} == {size}


@pytest.mark.parametrize("hw", [(100, 120), (256, 256), (321, 481), (130, 97), (77, 203)])
def test_plan_covers_every_output_pixel_exactly_once(hw):
    h, w = hw
    cover = np.zeros((h, w), dtype=np.int32)
    for lc in chop.plan(h, w):
        ty, tx, th, tw = lc.tile
        sy0, sy1, sx0, sx1 = lc.src
        dy0, dy1, dx0, dx1 = lc.dst
        assert 0 <= ty and ty + th <= h and 0 <= tx and tx + tw <= w
        assert 0 <= sy0 < sy1 <= th and 0 <= sx0 < sx1 <= tw
        assert (sy1 - sy0, sx1 - sx0) == (dy1 - dy0, dx1 - dx0)
        # stitching is a plain copy: the tile pixel and the output pixel are the same image pixel
        assert ty + sy0 == dy0 and tx + sx0 == dx0
        cover[dy0:dy1, dx0:dx1] += 1
    assert cover.min() == 1 and cover.max() == 1


def test_plan_rejects_images_too_small_to_chop():
    with pytest.raises(ValueError):
        chop.plan(30, 41)


def test_forward_chop_matches_reference_wrapper():
    z = load_npz("chop_probe.npz")
    for i in range(4):
        x, yref = z[f"x{i}"], z[f"y{i}"]
        for tb in (None, 5, 64):
            y = chop.forward_chop(probe_net, x, tile_batch=tb, distributed=False)
            assert torch.equal(y, yref), (i, tb, float((y - yref).abs().max()))


def test_x8_matches_reference_and_numpy_semantics():
    z = load_npz("chop_probe.npz")

    def net(t):
        return (t * torch.linspace(0.5, 1.5, t.shape[-1]).view(1, 1, 1, -1)
                + torch.linspace(-1, 1, t.shape[-2]).view(1, 1, -1, 1))

    y = chop.forward_x8(net, z["x8_in"])
    assert torch.allclose(y, z["x8_out"], rtol=0, atol=1e-6)
    # the 8 modes against numpy's flipud / rot90 on an (H, W, C, B) array (model/__init__.py:18-51)
    x = torch.arange(2 * 3 * 4 * 5, dtype=torch.float32).view(2, 3, 4, 5)
    a = x.numpy().transpose(2, 3, 1, 0)
    want = [a, np.flipud(np.rot90(a)), np.flipud(a), np.rot90(a, k=3), np.flipud(np.rot90(a, k=2)), np.rot90(a),
            np.rot90(a, k=2), np.flipud(np.rot90(a, k=3))]
    for mode in range(8):
        got = chop.augment(x, mode).numpy().transpose(2, 3, 1, 0)
        assert np.array_equal(got, want[mode]), mode
        inv = 8 - mode if mode in (3, 5) else mode
        assert torch.equal(chop.augment(chop.augment(x, mode), inv), x)


def test_chop_with_ensemble_matches_reference_wrapper():
    z = load_npz("chop_probe.npz")
    y = chop.forward_chop(probe_net, z["xe"], ensemble=True, distributed=False)
    assert torch.allclose(y, z["ye"], rtol=0, atol=2e-5)       # mean over 8 in a different association order


# ---- world_size-2 gloo: tiles sharded over ranks, same stitched image on every rank ----------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = torch.rand(1, 1, 100, 120, generator=torch.Generator().manual_seed(3))
        seen = []

        def net(t):
            seen.append(t.shape[0])
            return probe_net(t)

        y = chop.forward_chop(net, x, distributed=True)
        ref = chop.forward_chop(probe_net, x, distributed=False)
        q.put((rank, bool(torch.equal(y, ref)), sum(seen)))
    finally:
        dist.destroy_process_group()


def test_forward_chop_sharded_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    ntiles = len({lc.tile for lc in chop.plan(100, 120)})
    assert all(ok for _, ok, _ in res)
    assert sum(n for _, _, n in res) == ntiles and all(n > 0 for _, _, n in res)   # each rank ran its share only


# ---- GPU: the scheduler feeding the CUDA graph module with a batch of tiles -----------------------
@pytest.mark.gpu
def test_chop_over_cuda_ces_equals_tile_by_tile():
    """All leaf tiles in one batch through dagl_b200.CES == the reference's order (one tile at a time)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import dagl_b200
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    ces = dagl_b200.CES(in_channels=64).to(dev).eval()
    x = torch.randn(1, 64, 100, 120, device=dev)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            y_batched = chop.forward_chop(ces, x, tile_batch=64, distributed=False)
            y_single = chop.forward_chop(ces, x, tile_batch=1, distributed=False)
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    assert y_batched.shape == x.shape
    # The graph heads are exactly batch-independent (tests/test_ce_gpu.py::test_batch_independence_and_determinism), but
    # cuDNN picks other algorithms for the ResBlock / 1x1 convs at another batch size; those ~1e-6 differences move a few
    # threshold-tie neighbours in the following heads, i.e. the usual CES-level bar applies.
    assert (y_batched - y_single).abs().max().item() <= 4e-3 * y_single.abs().max().item()


# ---- Demosaic / DN_Real wrapper variants (Demosaic/model/__init__.py:107-114, 179-235, 265-299) ------------------
def test_demosaic_wrapper_shave12_matches_reference():
    """forward_chop with shave_size_max = 12 against outputs of the UNMODIFIED Demosaic wrapper
    (oracle/make_golden_chop_demosaic.py), bit-exact incl. a batch of two 3-colour images."""
    z = load_npz("chop_probe_demosaic.npz")
    for i in range(2):
        x, yref = z[f"x{i}"], z[f"y{i}"]
        for tb in (None, 3):
            y = chop.forward_chop(probe_net, x, shave_size_max=12, tile_batch=tb, distributed=False)
            assert torch.equal(y, yref), (i, tb, float((y - yref).abs().max()))
        # the DN_Gray value gives a different tiling: the parameter is what distinguishes the two wrappers
        assert not torch.equal(chop.forward_chop(probe_net, x, shave_size_max=24, distributed=False), yref)


def test_demosaic_wrapper_self_ensemble_matches_reference():
    """The wrapper's own 8-fold ensemble (Model.forward_x8: flip / flip / transpose lists, mean over the batch axis) around
    forward_chop, and around the bare network."""
    z = load_npz("chop_probe_demosaic.npz")
    x = z["xe"]
    y = chop.forward_x8_flips(lambda t: chop.forward_chop(probe_net, t, shave_size_max=12, distributed=False), x)
    assert y.shape == z["ye"].shape
    assert torch.allclose(y, z["ye"], rtol=0, atol=2e-6 * float(z["ye"].abs().max()))    # mean of 8: summation order only
    y2 = chop.forward_x8_flips(probe_net, x[:, :, :40, :52])
    assert torch.allclose(y2, z["ye_direct"], rtol=0, atol=2e-6 * float(z["ye_direct"].abs().max()))
