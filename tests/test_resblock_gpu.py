"""GPU parity of the ResBlock chain kernel (``dagl_resblocks_forward_f32``, through the ctypes binding) against the oracle
(oracle/resblock_oracle.py: the reference's ``common.ResBlock`` math, common.py:59-79, pinned to the unmodified class).

Floating point: the kernel sums the 576 products of an output in a different order than the CPU (tensor-core fp32
accumulation of fp16 hi/lo split operands, three terms) -- the bar is a relative error (max |d| / max |ref|) of
CHAIN_TOL = 3e-6 for chains of up to four blocks, i.e. fp32 re-association level (cuDNN's own fp32 kernels measure
0.4-1.8e-6 on the same cases, its TF32 default 1-2e-4).
"""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import resblock_oracle as RB
from oracle import ref_loader as R

pytestmark = pytest.mark.gpu
CHAIN_TOL = 3e-6
MODES = ("pair", "single", "auto")


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def rel_err(y, yref):
    return (y.double() - yref.double()).abs().max().item() / max(yref.double().abs().max().item(), 1e-30)


def make_blocks(params_list, dev, res_scale=1.0):
    import dagl_b200
    blocks = []
    for p in params_list:
        b = dagl_b200.ResBlock(64, res_scale=res_scale)
        if p["body.1.weight"].numel() != 1:
            b.body[1] = nn.PReLU(p["body.1.weight"].numel())
        if "body.0.bias" not in p:
            b.body[0].bias = None
            b.body[2].bias = None
        b.load_state_dict(p)
        blocks.append(b.to(dev).eval())
    return blocks


@pytest.mark.parametrize("mode", MODES)
def test_golden_chains(dev, mode):
    """The committed reference-made fixture (tests/golden/resblock_chain.npz)."""
    from dagl_b200.resblock import resblocks_forward
    from test_oracle import _resblock_golden
    for tag, params, x, y, rs in _resblock_golden():
        blocks = make_blocks(params, dev, rs)
        with torch.no_grad():
            got = resblocks_forward(blocks, x.to(dev), mode).cpu()
        assert rel_err(got, y) <= CHAIN_TOL, (tag, mode, rel_err(got, y))


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("shape,nb", [((1, 64, 64, 64), 1), ((2, 64, 37, 41), 2), ((1, 64, 5, 3), 1), ((1, 64, 1, 1), 2),
                                      ((3, 64, 72, 72), 4), ((1, 64, 3, 300), 1), ((1, 64, 131, 2), 1), ((1, 64, 128, 122), 1),
                                      # strip tiles (rows walked down 128-wide strips, one new halo row per tile): ragged second
                                      # strip, odd run counts (pair kernel: a dummy run), short / tall images, several per batch
                                      ((2, 64, 40, 250), 2), ((1, 64, 9, 104), 1), ((3, 64, 17, 128), 2), ((1, 64, 33, 256), 1),
                                      ((1, 64, 70, 384), 1)])
def test_chain_vs_oracle(dev, mode, shape, nb):
    """Seeded chains at ragged / tiny / chop-leaf shapes (tile boundaries inside rows, single-tile images, one-pixel images,
    W = 122: padded pitch exactly 128; widths that fill 128-wide strips run the row-reuse decomposition)."""
    from dagl_b200.resblock import resblocks_forward
    params = [RB.init_resblock_params(31 * nb + i) for i in range(nb)]
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(shape[2] * 1000 + shape[3]))
    want = RB.chain_forward(params, x)
    with torch.no_grad():
        got = resblocks_forward(make_blocks(params, dev), x.to(dev), mode).cpu()
    assert torch.isfinite(got).all()
    assert rel_err(got, want) <= CHAIN_TOL, rel_err(got, want)


def test_full_size_256_and_determinism(dev):
    from dagl_b200.resblock import resblocks_forward
    params = [RB.init_resblock_params(500 + i) for i in range(4)]
    x = torch.randn(1, 64, 256, 256, generator=torch.Generator().manual_seed(5))
    want = RB.chain_forward(params, x)
    blocks = make_blocks(params, dev)
    xs = x.to(dev)
    with torch.no_grad():
        a = resblocks_forward(blocks, xs, "pair")
        b = resblocks_forward(blocks, xs, "pair")
        c = resblocks_forward(blocks, xs, "single")
    assert torch.equal(a, b)                                   # no atomics on the data path: run-to-run identical
    assert rel_err(a.cpu(), want) <= CHAIN_TOL and rel_err(c.cpu(), want) <= CHAIN_TOL


def test_variants_per_channel_prelu_no_bias_res_scale(dev):
    from dagl_b200.resblock import resblocks_forward
    x = torch.randn(2, 64, 20, 23, generator=torch.Generator().manual_seed(9))
    for prelu_n, bias, rs in ((64, True, 1.0), (1, False, 1.0), (1, True, 0.1), (64, False, 2.0)):
        params = [RB.init_resblock_params(900 + i, prelu_n, bias) for i in range(2)]
        if prelu_n == 64:
            params[0]["body.1.weight"][::2] *= -3.0              # |slope| > 1 and negative slopes: the bound must cover them
        want = RB.chain_forward(params, x, rs)
        with torch.no_grad():
            got = resblocks_forward(make_blocks(params, dev, rs), x.to(dev), "pair").cpu()
        assert rel_err(got, want) <= CHAIN_TOL, (prelu_n, bias, rs, rel_err(got, want))


@pytest.mark.parametrize("scale", [1e-6, 1.0, 3e4, 1e12])
def test_input_magnitudes(dev, scale):
    """The fp16 operand images are scaled per image from measured maxima: tiny and huge activations keep fp32 accuracy.
    Bias-free blocks are positively homogeneous, so the relative error must not depend on the scale."""
    from dagl_b200.resblock import resblocks_forward
    params = [RB.init_resblock_params(40 + i, 1, False) for i in range(3)]
    x = torch.randn(2, 64, 33, 29, generator=torch.Generator().manual_seed(4)) * scale
    x[1] *= 1e-3                                               # images of one batch with very different ranges
    want = RB.chain_forward(params, x)
    with torch.no_grad():
        got = resblocks_forward(make_blocks(params, dev), x.to(dev), "pair").cpu()
    assert torch.isfinite(got).all()
    for i in range(2):
        assert rel_err(got[i], want[i]) <= CHAIN_TOL, (scale, i, rel_err(got[i], want[i]))


def test_zero_input_and_inplace_output(dev):
    from dagl_b200 import _lib
    from dagl_b200.resblock import resblocks_forward, _struct
    params = [RB.init_resblock_params(77)]
    blocks = make_blocks(params, dev)
    z = torch.zeros(1, 64, 16, 16, device=dev)
    with torch.no_grad():
        got = resblocks_forward(blocks, z, "pair").cpu()
    assert rel_err(got, RB.chain_forward(params, z.cpu())) <= CHAIN_TOL      # only the biases propagate
    # y may alias x (include/dagl_b200.h): call the C-ABI directly with y == x
    L = _lib.lib()
    x = torch.randn(1, 64, 24, 20, generator=torch.Generator().manual_seed(8)).to(dev)
    want = RB.chain_forward(params, x.cpu())
    ws = torch.empty(L.dagl_resblocks_workspace_bytes(1, 1, 64, 24, 20), dtype=torch.uint8, device=dev)
    arr = (_lib.DaglResBlockWeights * 1)(_struct(blocks[0], None))
    rc = L.dagl_resblocks_forward_f32(arr, 1, x.data_ptr(), x.data_ptr(), 1, 64, 24, 20, ws.data_ptr(), ws.numel(), 0,
                                      torch.cuda.current_stream().cuda_stream)
    assert rc == 0, L.dagl_last_error()
    torch.cuda.synchronize()
    assert rel_err(x.cpu(), want) <= CHAIN_TOL


def test_error_codes(dev):
    from dagl_b200 import _lib
    from dagl_b200.resblock import _struct
    L = _lib.lib()
    blocks = make_blocks([RB.init_resblock_params(1)], dev)
    arr = (_lib.DaglResBlockWeights * 1)(_struct(blocks[0], None))
    x = torch.zeros(1, 64, 8, 8, device=dev)
    ws = torch.empty(L.dagl_resblocks_workspace_bytes(1, 1, 64, 8, 8), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    assert L.dagl_resblocks_workspace_bytes(1, 1, 32, 8, 8) == 0
    assert L.dagl_resblocks_forward_f32(arr, 1, x.data_ptr(), x.data_ptr(), 1, 32, 8, 8, ws.data_ptr(), ws.numel(), 0, st) == -2
    assert L.dagl_resblocks_forward_f32(arr, 1, x.data_ptr(), x.data_ptr(), 1, 64, 8, 8, ws.data_ptr(), 1024, 0, st) == -3
    assert L.dagl_resblocks_forward_f32(arr, 1, None, x.data_ptr(), 1, 64, 8, 8, ws.data_ptr(), ws.numel(), 0, st) == -1
    assert L.dagl_resblocks_forward_f32(arr, 1, x.data_ptr(), x.data_ptr(), 1, 64, 8, 8, ws.data_ptr(), ws.numel(), 7, st) == -2


def test_fused_sequential_and_training_path(dev):
    """CES.RBS1 of dagl_b200.CES runs the chain kernel under no_grad and the torch modules when a gradient is needed."""
    import dagl_b200
    torch.manual_seed(3)
    ces = dagl_b200.CES(64).to(dev).eval()
    x = torch.randn(1, 64, 24, 24, device=dev)
    params = [{k: v.detach().cpu() for k, v in blk.state_dict().items()} for blk in ces.RBS1]
    want = RB.chain_forward(params, x.cpu())
    L = dagl_b200._lib.lib()
    with torch.no_grad():
        got = ces.RBS1(x)
    assert L.dagl_last_impl().decode().startswith("resblock_tc")
    assert rel_err(got.cpu(), want) <= CHAIN_TOL
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        xg = x.clone().requires_grad_(True)
        y = ces.RBS1(xg)
        y.sum().backward()
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    assert xg.grad is not None and torch.isfinite(xg.grad).all()
    assert rel_err(y.detach().cpu(), want) <= 1e-5


def test_trained_resblocks_from_checkpoint(dev):
    """The sixteen RR.body ResBlocks + the eight of CES with the reference's shipped DN_Gray weights, on a real feature map
    (head conv of a seeded image): two chains of eight and two of four, vs the oracle."""
    ck = R.checkpoint("DN_Gray") if R.available("DN_Gray") else None
    if ck is None:
        pytest.skip("baseline/_ref/DN_Gray checkpoint not present")
    from dagl_b200.resblock import resblocks_forward
    sd = torch.load(ck, map_location="cpu")
    img = torch.rand(1, 1, 96, 80, generator=torch.Generator().manual_seed(2))
    feat = torch.nn.functional.conv2d(img, sd["head.0.weight"], sd["head.0.bias"], padding=1)
    def chain(prefixes):
        return [{k: sd[f"{p}.{k}"] for k in ("body.0.weight", "body.0.bias", "body.1.weight", "body.2.weight", "body.2.bias")}
                for p in prefixes]
    groups = {"body[0:8]": [f"body.{i}" for i in range(8)], "body[9:17]": [f"body.{i}" for i in range(9, 17)],
              "RBS1": [f"body.8.RBS1.{i}" for i in range(4)], "RBS2": [f"body.8.RBS2.{i}" for i in range(4)]}
    for name, pre in groups.items():
        params = chain(pre)
        want = RB.chain_forward(params, feat)
        with torch.no_grad():
            got = resblocks_forward(make_blocks(params, dev), feat.to(dev), "pair").cpu()
        e = rel_err(got, want)
        assert e <= 2 * CHAIN_TOL, (name, e)                   # chains of eight: twice the four-block budget


def test_repeated_calls_overwrite_poisoned_outputs(dev):
    """Back-to-back chain calls (no host synchronisation in between, alternating shapes / kernels, the output buffer poisoned
    with NaN before every call): every call must fully overwrite its output with the same bits."""
    from dagl_b200 import _lib
    from dagl_b200.resblock import _struct
    L = _lib.lib()
    params = [RB.init_resblock_params(60 + i) for i in range(2)]
    blocks = make_blocks(params, dev)
    arr = (_lib.DaglResBlockWeights * 2)(*[_struct(b, None) for b in blocks])
    st = torch.cuda.current_stream().cuda_stream
    cases = []
    for shape, mode in (((1, 64, 64, 72), 0), ((2, 64, 24, 128), 1), ((1, 64, 40, 250), 0), ((3, 64, 30, 30), 2)):
        x = torch.randn(*shape, generator=torch.Generator().manual_seed(shape[3])).to(dev)
        ws = torch.empty(L.dagl_resblocks_workspace_bytes(2, *shape), dtype=torch.uint8, device=dev)
        cases.append((shape, mode, x, ws, torch.empty_like(x), RB.chain_forward(params, x.cpu())))
    first = {}
    for rep in range(12):
        for i, (shape, mode, x, ws, y, want) in enumerate(cases):
            y.fill_(float("nan"))
            rc = L.dagl_resblocks_forward_f32(arr, 2, x.data_ptr(), y.data_ptr(), *shape, ws.data_ptr(), ws.numel(), mode, st)
            assert rc == 0, L.dagl_last_error()
            if rep == 0:
                first[i] = y.clone()
                assert rel_err(first[i].cpu(), want) <= CHAIN_TOL
            else:
                assert torch.equal(y, first[i]), (rep, shape, mode)


def test_outputs_are_ordered_before_following_torch_ops(dev):
    """The chain's kernels are chained by programmatic dependent launch; whatever torch enqueues next on the stream must see
    the finished output.  200 calls into freshly allocated (NaN-poisoned) outputs, each consumed at once by a torch kernel
    with no host synchronisation, at a launch-bound and at a longer shape, single-block and four-block chains."""
    from dagl_b200.resblock import resblocks_forward
    params = [RB.init_resblock_params(70 + i) for i in range(4)]
    blocks = make_blocks(params, dev)
    for shape, nb in (((1, 64, 16, 16), 1), ((1, 64, 64, 64), 4), ((1, 64, 256, 256), 4), ((8, 64, 72, 72), 2)):
        x = torch.randn(*shape, generator=torch.Generator().manual_seed(3)).to(dev)
        with torch.no_grad():
            first = resblocks_forward(blocks[:nb], x).clone()
            bad = torch.zeros((), device=dev)
            for i in range(50):
                for _ in range(3):
                    torch.empty_like(x).fill_(float("nan"))            # poison the free blocks the next output will get
                y = resblocks_forward(blocks[:nb], x)
                d = (y - first).abs().max()                             # consumed immediately, no synchronisation
                bad = torch.maximum(bad, torch.nan_to_num(d, nan=1e30))
                del y
        assert float(bad) == 0.0, (shape, nb, float(bad))
