"""CPU tests of the ResBlock-chain host logic (no GPU): which modules are taken over, that containers keep working on CPU
tensors, the C-ABI's host-side argument checks, and that ``dagl_b200.RR`` is state-dict compatible with the reference RR."""
import ctypes as C

import pytest
import torch
import torch.nn as nn

import dagl_b200
from dagl_b200 import _lib
from dagl_b200 import resblock as DR
from oracle import ref_loader as R
from oracle import resblock_oracle as RB


def test_only_reference_shaped_resblocks_are_taken_over():
    ok = dagl_b200.ResBlock(64)
    assert DR.is_resblock(ok)
    assert not DR.is_resblock(dagl_b200.ResBlock(32))                       # the kernel is built for 64 channels
    strided = dagl_b200.ResBlock(64); strided.body[0] = nn.Conv2d(64, 64, 3, stride=2, padding=1)
    assert not DR.is_resblock(strided)
    relu = dagl_b200.ResBlock(64); relu.body[1] = nn.ReLU()
    assert not DR.is_resblock(relu)
    half = dagl_b200.ResBlock(64).half()
    assert not DR.is_resblock(half)
    four = dagl_b200.ResBlock(64); four.body = nn.Sequential(*four.body, nn.Identity())   # e.g. the bn=True layout
    assert not DR.is_resblock(four)
    assert not DR.is_resblock(nn.Conv2d(64, 64, 3))


def test_fuse_sequential_is_in_place_idempotent_and_cpu_safe():
    torch.manual_seed(0)
    blocks = [dagl_b200.ResBlock(64) for _ in range(3)]
    seq = nn.Sequential(blocks[0], blocks[1], nn.Conv2d(64, 64, 1), blocks[2]).eval()
    plain = nn.Sequential(nn.Conv2d(64, 64, 1), nn.ReLU())
    assert not DR.fuse_sequential(plain) and type(plain) is nn.Sequential
    keys = list(seq.state_dict())
    assert DR.fuse_sequential(seq) and DR.is_fused(seq) and isinstance(seq, nn.Sequential)
    cls = type(seq)
    assert DR.fuse_sequential(seq) and type(seq) is cls                     # idempotent
    assert list(seq.state_dict()) == keys
    x = torch.randn(2, 64, 9, 7)
    with torch.no_grad():
        want = x
        for m in seq:
            want = m(want)
        got = seq(x)                                                        # CPU tensor: the blocks' own torch modules
    assert torch.equal(got, want)
    params = [{k: v for k, v in b.state_dict().items()} for b in blocks[:2]]
    assert torch.equal(RB.chain_forward(params, x), blocks[1](blocks[0](x)))   # and those are the reference's math


def test_resblocks_forward_has_no_cpu_path():
    with pytest.raises(RuntimeError, match="no CPU path"):
        DR.resblocks_forward([dagl_b200.ResBlock(64)], torch.zeros(1, 64, 4, 4))
    with pytest.raises(RuntimeError, match="no CPU path"):
        dagl_b200.RR().eval()(torch.zeros(1, 1, 8, 8))                      # the graph heads have none either


def test_c_abi_host_side_checks():
    L = _lib.lib()
    assert L.dagl_resblock_packed_weights_bytes() >= 2 * 2 * 36 * 2048
    small, big = L.dagl_resblocks_workspace_bytes(4, 1, 64, 64, 64), L.dagl_resblocks_workspace_bytes(4, 1, 64, 256, 256)
    assert 0 < small < big
    assert L.dagl_resblocks_workspace_bytes(1, 1, 64, 64, 64) < small       # a single block needs no fp32 ping-pong buffers
    assert L.dagl_resblocks_workspace_bytes(4, 1, 32, 64, 64) == 0 and L.dagl_resblocks_workspace_bytes(0, 1, 64, 64, 64) == 0
    fake = C.c_void_p(16)
    w = _lib.DaglResBlockWeights(16, 16, 16, 1, 16, 16, 1.0, None)
    arr = (_lib.DaglResBlockWeights * 1)(w)
    assert L.dagl_resblocks_forward_f32(arr, 1, None, fake, 1, 64, 8, 8, fake, 1 << 30, 0, None) == -1
    assert L.dagl_resblocks_forward_f32(arr, 1, fake, fake, 1, 48, 8, 8, fake, 1 << 30, 0, None) == -2
    assert L.dagl_resblocks_forward_f32(arr, 1, fake, fake, 1, 64, 8, 8, fake, 1 << 30, 5, None) == -2
    bad = (_lib.DaglResBlockWeights * 1)(_lib.DaglResBlockWeights(16, 16, 16, 3, 16, 16, 1.0, None))
    assert L.dagl_resblocks_forward_f32(bad, 1, fake, fake, 1, 64, 8, 8, fake, 1 << 30, 0, None) == -1
    assert b"PReLU" in L.dagl_last_error()
    assert L.dagl_resblock_pack_weights_f32(C.byref(w), None, 0, None) == -1


@pytest.mark.skipif(not R.available("DN_Gray"), reason="reference sources not present")
@pytest.mark.parametrize("task", ["DN_Gray", "CAR", "Demosaic"])
def test_rr_mirror_state_dict_matches_reference(task):
    """dagl_b200.RR has the reference RR's parameter names, shapes and (for the constant add_mean) values; the shipped
    checkpoint loads with strict=True."""
    if not R.available(task):
        pytest.skip(f"{task} not present")
    ref = R.load_task(task)
    args = R.rr_args(task)
    torch.manual_seed(0)
    net = ref.dagl.RR(args)
    nrb = 32 if task == "Demosaic" else args.n_resblocks          # Demosaic/model/dagl.py:14 hard-codes 32 (16 + CES + 16)
    mine = dagl_b200.RR(n_resblocks=nrb, n_feats=args.n_feats, n_colors=args.n_colors,
                        res_scale=args.res_scale, rgb_range=args.rgb_range)
    a, b = net.state_dict(), mine.state_dict()
    assert list(a) == list(b)
    assert all(a[k].shape == b[k].shape for k in a)
    assert torch.equal(a["add_mean.weight"], b["add_mean.weight"]) and torch.equal(a["add_mean.bias"], b["add_mean.bias"])
    ck = R.checkpoint(task)
    if ck is not None:
        mine.load_state_dict(torch.load(ck, map_location="cpu"), strict=True)
    assert sum(DR.is_resblock(m) for m in mine.modules()) == nrb + 8 and DR.is_fused(mine.body)
